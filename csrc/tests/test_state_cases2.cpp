// The rest of the reference's state suite, one case per case / section
// (reference: tests/test/state/test_state.cpp:382-396, 668-712, 915-1008,
// tests/test/state/test_state_server.cpp:61-222)
#include "harness.h"

#include <faabric/redis/Redis.h>
#include <faabric/state/InMemoryStateKeyValue.h>
#include <faabric/state/State.h>
#include <faabric/state/StateClient.h>
#include <faabric/state/StateServer.h>
#include <faabric/util/config.h>
#include <faabric/util/memory.h>

#include <cstring>
#include <sys/mman.h>

using namespace faabric::state;

namespace {
const size_t P = faabric::util::HOST_PAGE_SIZE;

// This host holds the main copies and serves them; `remoteState` is how
// another host sees them
struct Hosts
{
    State& mainState = getGlobalState();
    StateServer server;
    State remoteState;

    Hosts()
      : server(getGlobalState())
      , remoteState("otherhost")
    {
        faabric::util::getSystemConfig().reset();
        mainState.forceClearAll(true);
        getInMemoryStateRegistry().clear();
        faabric::redis::Redis::getState().flushAll();
        server.start();
    }

    ~Hosts()
    {
        server.stop();
        mainState.forceClearAll(true);
        remoteState.forceClearAll(false);
        getInMemoryStateRegistry().clear();
        faabric::util::getSystemConfig().reset();
    }

    std::shared_ptr<InMemoryStateKeyValue> mainKv(const std::string& user, const std::string& key, size_t size)
    {
        return std::static_pointer_cast<InMemoryStateKeyValue>(mainState.getKV(user, key, size));
    }
};

const std::vector<uint8_t> dataA = { 0, 1, 2, 3, 4, 5, 6, 7 };
const std::vector<uint8_t> dataB = { 7, 6, 5, 4, 3, 2, 1, 0 };
}

TEST_CASE("state case: a chunk may not reach past the end of the allocated storage", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainState.getKV("cases", "oversize", 2);
    std::vector<uint8_t> update = { 8, 8, 8 };
    REQUIRE_THROWS(kv->setChunk((long)P - 2, update.data(), 3));
    // (the bound is the page-rounded storage, as in the reference)
    kv->setChunk(0, update.data(), 2);
}

TEST_CASE("state case: mappings of pages the replica has not pulled yet", "[state][cases]")
{
    Hosts h;
    const size_t mappingSize = 3 * P;
    const size_t totalSize = 10 * P + 15;
    std::vector<uint8_t> values(totalSize, 1);
    auto mainKv = h.mainState.getKV("cases", "uninit-map", totalSize);
    mainKv->set(values.data());
    mainKv->pushFull();

    void* regionA = ::mmap(nullptr, mappingSize, PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    void* regionB = ::mmap(nullptr, mappingSize, PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    auto localKv = h.remoteState.getKV("cases", "uninit-map", totalSize);
    localKv->mapSharedMemory(regionA, 6, 3);
    localKv->mapSharedMemory(regionB, 2, 3);
    auto* bytesA = (uint8_t*)regionA;
    auto* bytesB = (uint8_t*)regionB;
    // direct pointers pull implicitly
    uint8_t* chunkA = localKv->getChunk(6 * P, 10);
    uint8_t* chunkB = localKv->getChunk(2 * P, 10);
    bytesA[5] = 5;
    bytesB[9] = 9;
    REQUIRE_EQ((int)chunkA[0], 1);
    REQUIRE_EQ((int)chunkB[0], 1);
    REQUIRE_EQ((int)chunkA[5], 5);
    REQUIRE_EQ((int)chunkB[9], 9);
    localKv->unmapSharedMemory(regionA, 3);
    localKv->unmapSharedMemory(regionB, 3);
}

TEST_CASE("state case: two disjoint chunks from one page are pulled one after the other", "[state][cases]")
{
    Hosts h;
    const size_t valueSize = 20 * P + 123;
    std::vector<uint8_t> values(valueSize, 1);
    auto mainKv = h.mainState.getKV("cases", "disjoint", valueSize);
    mainKv->set(values.data());
    mainKv->pushFull();

    const long offsetA = 2 * P + 10, offsetB = 2 * P + 20, lenA = 5, lenB = 10;
    std::vector<uint8_t> actualA(lenA, 0), actualB(lenB, 0);
    auto localKv = h.remoteState.getKV("cases", "disjoint", valueSize);
    localKv->getChunk(offsetA, actualA.data(), lenA);
    localKv->getChunk(offsetB, actualB.data(), lenB);
    // storage is reserved in whole pages
    REQUIRE_EQ(localKv->getSharedMemorySize(), 21 * P);
    REQUIRE(actualA == std::vector<uint8_t>(lenA, 1));
    REQUIRE(actualB == std::vector<uint8_t>(lenB, 1));
}

TEST_CASE("state case: a replica writes locally, the main copy changes on push", "[state][cases]")
{
    Hosts h;
    REQUIRE_EQ(h.remoteState.getKVCount(), 0u);
    auto mainKv = h.mainKv("foo", "bar", dataA.size());
    mainKv->set(dataA.data());
    mainKv->pushFull();
    // the size is known before the value is touched on the replica
    REQUIRE_EQ(h.remoteState.getStateSize("foo", "bar"), dataA.size());
    auto localKv = std::static_pointer_cast<InMemoryStateKeyValue>(h.remoteState.getKV("foo", "bar", dataA.size()));
    REQUIRE(!localKv->isMaster());
    REQUIRE(mainKv->isMaster());
    localKv->set(dataB.data());
    std::vector<uint8_t> actualLocal(dataA.size(), 0);
    localKv->get(actualLocal.data());
    REQUIRE(actualLocal == dataB);
    std::vector<uint8_t> actualMain(dataA.size(), 0);
    mainKv->get(actualMain.data());
    REQUIRE(actualMain == dataA);
    localKv->pushFull();
    mainKv->get(actualMain.data());
    REQUIRE(actualMain == dataB);
}

TEST_CASE("state server case: thread count from the config", "[state][cases]")
{
    auto& conf = faabric::util::getSystemConfig();
    conf.stateServerThreads = 7;
    {
        StateServer server(getGlobalState());
        REQUIRE_EQ(server.getNThreads(), 7);
    }
    conf.reset();
}

TEST_CASE("state server case: size request", "[state][cases]")
{
    Hosts h;
    h.mainKv("foo", "bar", dataA.size())->set(dataA.data());
    StateClient client("foo", "bar", DEFAULT_STATE_HOST);
    REQUIRE_EQ(client.stateSize(), dataA.size());
}

TEST_CASE("state server case: pulling several overlapping chunks in one request", "[state][cases]")
{
    Hosts h;
    h.mainKv("foo", "bar", dataA.size())->set(dataA.data());
    StateClient client("foo", "bar", DEFAULT_STATE_HOST);
    std::vector<StateChunk> chunks = { StateChunk(1, 3, nullptr), StateChunk(2, 4, nullptr), StateChunk(7, 1, nullptr) };
    std::vector<uint8_t> actual(dataA.size(), 0);
    client.pullChunks(chunks, actual.data());
    REQUIRE(actual == (std::vector<uint8_t>{ 0, 1, 2, 3, 4, 5, 0, 7 }));
}

TEST_CASE("state server case: pushing several overlapping chunks, applied in order", "[state][cases]")
{
    Hosts h;
    auto kvA = h.mainKv("foo", "bar", dataA.size());
    kvA->set(dataA.data());
    StateClient client("foo", "bar", DEFAULT_STATE_HOST);
    std::vector<uint8_t> a = { 7, 7 }, b = { 8 }, c = { 9, 9, 9 };
    std::vector<StateChunk> chunks = { StateChunk(0, a), StateChunk(6, b), StateChunk(1, c) };
    client.pushChunks(chunks);
    std::vector<uint8_t> actual(dataA.size(), 0);
    kvA->get(actual.data());
    REQUIRE(actual == (std::vector<uint8_t>{ 7, 9, 9, 9, 4, 5, 8, 7 }));
}

TEST_CASE("state server case: appends through the client, read back in order", "[state][cases]")
{
    Hosts h;
    h.mainKv("foo", "bar", dataA.size())->set(dataA.data());
    StateClient client("foo", "bar", DEFAULT_STATE_HOST);
    std::vector<uint8_t> a = { 3, 2, 1 }, b = { 5, 5 }, c = { 2, 2 };
    client.append(a.data(), a.size());
    client.append(b.data(), b.size());
    client.append(c.data(), c.size());
    std::vector<uint8_t> actual(7, 0);
    client.pullAppended(actual.data(), actual.size(), 3);
    REQUIRE(actual == (std::vector<uint8_t>{ 3, 2, 1, 5, 5, 2, 2 }));
    client.clearAppended();
    REQUIRE_THROWS(client.pullAppended(actual.data(), actual.size(), 3));
}

TEST_CASE("state server case: push and size on the main host itself", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainKv("foo", "bar", dataA.size());
    REQUIRE(kv->isMaster());
    kv->set(dataA.data());
    kv->pushFull(); // nothing to send, nothing breaks
    REQUIRE_EQ(h.mainState.getStateSize("foo", "bar"), dataA.size());
}

TEST_CASE("state server case: appends on the main host itself", "[state][cases]")
{
    Hosts h;
    std::vector<uint8_t> a = { 1, 1 }, b = { 2, 2, 2 }, c = { 3, 3 };
    auto kv = h.mainKv("foo", "bar", 1); // appended values do not live in the value's storage
    kv->append(a.data(), a.size());
    kv->append(b.data(), b.size());
    kv->append(c.data(), c.size());
    std::vector<uint8_t> actual(7, 0);
    kv->getAppended(actual.data(), actual.size(), 3);
    REQUIRE(actual == (std::vector<uint8_t>{ 1, 1, 2, 2, 2, 3, 3 }));
}

TEST_CASE("state server case: a pull on the main host keeps what was set there", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainKv("foo", "bar", dataA.size());
    kv->set(dataA.data());
    REQUIRE(kv->isMaster());
    kv->pushFull();
    kv->set(dataB.data());
    auto again = h.mainState.getKV("foo", "bar", dataA.size());
    again->pull();
    std::vector<uint8_t> actual(again->get(), again->get() + dataA.size());
    REQUIRE(actual == dataB);
}
