// State values in redis mode, one case per case of the reference's suite,
// checked against the (emulated) store itself
// (reference: tests/test/state/test_redis_state.cpp:20-560)
#include "harness.h"

#include <faabric/redis/Redis.h>
#include <faabric/state/State.h>
#include <faabric/util/config.h>
#include <faabric/util/memory.h>
#include <faabric/util/state.h>

#include <sys/mman.h>

#include <cstring>

using namespace faabric::state;
using faabric::redis::Redis;

namespace {
struct RedisState
{
    State& state = getGlobalState();
    Redis& store = Redis::getState();
    int counter = 0;

    RedisState()
    {
        faabric::util::getSystemConfig().reset();
        state.forceClearAll(true);
        getInMemoryStateRegistry().clear();
        store.flushAll();
        faabric::util::getSystemConfig().stateMode = "redis";
    }

    ~RedisState()
    {
        state.forceClearAll(true);
        store.flushAll();
        faabric::util::getSystemConfig().reset();
    }

    std::shared_ptr<StateKeyValue> kvOf(size_t size) { return state.getKV("demo", "redis_case_" + std::to_string(counter++), size); }

    static std::string keyOf(const std::shared_ptr<StateKeyValue>& kv) { return faabric::util::keyForUser(kv->user, kv->key); }
};
}

TEST_CASE("redis state case: sizes, set or still unknown", "[state][redis][cases]")
{
    RedisState f;
    REQUIRE_EQ(f.state.getStateSize("demo", "nothing_here"), 0u);
    auto kv = f.kvOf(5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 };
    kv->set(v.data());
    kv->pushFull();
    REQUIRE_EQ(f.state.getStateSize(kv->user, kv->key), 5u);
    REQUIRE_EQ(kv->size(), 5u);
    // a size-less handle learns it from the store
    f.state.forceClearAll(false);
    auto again = f.state.getKV(kv->user, kv->key);
    std::vector<uint8_t> got(5);
    again->get(got.data()); // (the size is resolved on first use, as in the reference)
    REQUIRE_EQ(again->size(), 5u);
    REQUIRE(got == v);
}

TEST_CASE("redis state case: simple get and set", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 }, got(5);
    kv->set(v.data());
    kv->get(got.data());
    REQUIRE(got == v);
    // not in the store until pushed
    REQUIRE(f.store.get(f.keyOf(kv)).empty());
    kv->pushFull();
    REQUIRE(f.store.get(f.keyOf(kv)) == v);
    // a new value in memory, again only after the push in the store
    v = { 5, 5, 5, 5, 5 };
    kv->set(v.data());
    kv->get(got.data());
    REQUIRE(got == v);
    REQUIRE(f.store.get(f.keyOf(kv)) == (std::vector<uint8_t>{ 0, 1, 2, 3, 4 }));
    kv->pushFull();
    REQUIRE(f.store.get(f.keyOf(kv)) == v);
}

TEST_CASE("redis state case: get and set of a segment", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(10);
    std::vector<uint8_t> v = { 0, 0, 1, 1, 2, 2, 3, 3, 4, 4 };
    kv->set(v.data());
    kv->pushFull();
    std::vector<uint8_t> update = { 8, 8, 8 };
    kv->setChunk(6, update.data(), 3);
    std::vector<uint8_t> seg(3);
    kv->getChunk(6, seg.data(), 3);
    REQUIRE(seg == update);
    std::vector<uint8_t> whole(10);
    kv->get(whole.data());
    REQUIRE(whole == (std::vector<uint8_t>{ 0, 0, 1, 1, 2, 2, 8, 8, 8, 4 }));
    // the store changes with the partial push only
    REQUIRE(f.store.get(f.keyOf(kv)) == v);
    kv->pushPartial();
    REQUIRE(f.store.get(f.keyOf(kv)) == whole);
}

TEST_CASE("redis state case: reading a segment of a value dropped locally pulls it", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(10);
    std::vector<uint8_t> v = { 0, 0, 1, 1, 2, 2, 3, 3, 4, 4 };
    kv->set(v.data());
    kv->pushFull();
    std::string user = kv->user, key = kv->key;
    f.state.forceClearAll(false);
    REQUIRE(f.store.get(faabric::util::keyForUser(user, key)) == v);
    auto after = f.state.getKV(user, key, v.size());
    std::vector<uint8_t> seg(3, 0);
    after->getChunk(4, seg.data(), 3);
    REQUIRE(seg == (std::vector<uint8_t>{ 2, 2, 3 }));
}

TEST_CASE("redis state case: only segments marked dirty are pushed", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(10);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9 };
    kv->set(v.data());
    kv->pushFull();
    uint8_t* ptr = kv->get();
    ptr[0] = 8;
    ptr[5] = 7;
    kv->flagChunkDirty(0, 2);
    kv->pushPartial();
    v[0] = 8;
    REQUIRE(f.store.get(f.keyOf(kv)) == v); // byte 5 stayed behind
    // the in-memory value keeps both edits
    REQUIRE_EQ((int)ptr[5], 7);
}

TEST_CASE("redis state case: several dirty segments next to direct updates of the store", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(20);
    std::vector<uint8_t> zeros(20, 0);
    kv->set(zeros.data());
    kv->pushFull();
    uint8_t* p = kv->get();
    p[1] = 1;
    p[2] = 2;
    p[3] = 3;
    p[10] = 4;
    p[11] = 5;
    p[14] = p[15] = p[16] = p[17] = 7;
    kv->flagChunkDirty(1, 3);
    kv->flagChunkDirty(10, 2);
    kv->flagChunkDirty(14, 4);
    // somebody else writes to the store: next to and over our segments
    std::vector<uint8_t> directA = { 2, 2 }, directB = { 6, 6, 6, 6, 6 };
    f.store.setRange(f.keyOf(kv), 6, directA.data(), 2);
    f.store.setRange(f.keyOf(kv), 0, directB.data(), 5);
    kv->pushPartial();
    // our segments win where they overlap, everything else is kept
    REQUIRE(f.store.get(f.keyOf(kv)) ==
            (std::vector<uint8_t>{ 6, 1, 2, 3, 6, 0, 2, 2, 0, 0, 4, 5, 0, 0, 7, 7, 7, 7, 0, 0 }));
}

TEST_CASE("redis state case: partial update of doubles", "[state][redis][cases]")
{
    RedisState f;
    const long n = 20;
    auto kv = f.kvOf(n * sizeof(double));
    std::vector<double> expected(n, 0.0);
    std::vector<uint8_t> zeros(n * sizeof(double), 0);
    kv->set(zeros.data());
    kv->pushFull();
    auto* actual = reinterpret_cast<double*>(kv->get());
    for (auto [idx, val] : std::vector<std::pair<int, double>>{ { 0, 123.456 }, { 1, -100304.223 }, { 9, 6090293.222 }, { 13, -123.444 } }) {
        actual[idx] = val;
        expected[idx] = val;
        kv->flagChunkDirty(idx * sizeof(double), sizeof(double));
    }
    kv->pushPartial();
    auto* after = reinterpret_cast<double*>(kv->get());
    REQUIRE(std::vector<double>(after, after + n) == expected);
    std::vector<double> stored(n);
    f.store.get(f.keyOf(kv), BYTES(stored.data()), n * sizeof(double));
    REQUIRE(stored == expected);
}

TEST_CASE("redis state case: partial sets of just the first and the last element", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 };
    kv->set(v.data());
    kv->pushFull();
    std::vector<uint8_t> update = { 8 };
    kv->setChunk(4, update.data(), 1);
    kv->pushPartial();
    REQUIRE(f.store.get(f.keyOf(kv)) == (std::vector<uint8_t>{ 0, 1, 2, 3, 8 }));
    kv->setChunk(0, update.data(), 1);
    kv->pushPartial();
    kv->pushPartial(); // nothing dirty the second time
    REQUIRE(f.store.get(f.keyOf(kv)) == (std::vector<uint8_t>{ 8, 1, 2, 3, 8 }));
    update = { 6 };
    kv->setChunk(0, update.data(), 1);
    kv->setChunk(4, update.data(), 1);
    kv->pushPartial();
    REQUIRE(f.store.get(f.keyOf(kv)) == (std::vector<uint8_t>{ 6, 1, 2, 3, 6 }));
}

TEST_CASE("redis state case: push with a mask value", "[state][redis][cases]")
{
    RedisState f;
    const size_t size = 4 * sizeof(double);
    auto data = f.kvOf(size);
    auto mask = f.kvOf(size);
    auto* d = reinterpret_cast<double*>(data->get());
    std::vector<double> initial = { 1.2345, 12.345, 987.6543, 10987654.3 };
    std::copy(initial.begin(), initial.end(), d);
    data->flagDirty();
    data->pushFull();
    auto stored = f.store.get(f.keyOf(data));
    REQUIRE(std::vector<double>((double*)stored.data(), (double*)stored.data() + 4) == initial);
    d[1] = 11.11;
    d[2] = 222.222;
    d[3] = 3333.3333;
    auto* m = reinterpret_cast<unsigned int*>(mask->get());
    faabric::util::maskDouble(m, 1);
    faabric::util::maskDouble(m, 3);
    data->flagDirty();
    data->pushPartialMask(mask);
    stored = f.store.get(f.keyOf(data));
    // old, new (masked), old (changed in memory but not masked), new (masked)
    REQUIRE(std::vector<double>((double*)stored.data(), (double*)stored.data() + 4) ==
            (std::vector<double>{ 1.2345, 11.11, 987.6543, 3333.3333 }));
}

TEST_CASE("redis state case: a plain get never pulls behind the caller's back", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(4);
    std::vector<uint8_t> v = { 0, 1, 2, 3 };
    kv->set(v.data());
    kv->pushFull();
    REQUIRE(f.store.get(f.keyOf(kv)) == v);
    f.store.set(f.keyOf(kv), { 5, 5, 5, 5 });
    std::vector<uint8_t> got(4);
    kv->get(got.data());
    REQUIRE(got == v);
    // an explicit pull does
    kv->pull();
    kv->get(got.data());
    REQUIRE(got == (std::vector<uint8_t>{ 5, 5, 5, 5 }));
}

TEST_CASE("redis state case: a full push only happens when something is dirty", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(4);
    std::vector<uint8_t> v = { 0, 1, 2, 3 };
    kv->set(v.data());
    kv->pushFull();
    std::vector<uint8_t> direct = { 3, 4, 5, 6 };
    f.store.set(f.keyOf(kv), direct);
    kv->pushFull(); // clean: the store keeps what was put there directly
    REQUIRE(f.store.get(f.keyOf(kv)) == direct);
    std::vector<uint8_t> v2 = { 7, 7, 7, 7 };
    kv->set(v2.data());
    kv->pushFull();
    REQUIRE(f.store.get(f.keyOf(kv)) == v2);
}

TEST_CASE("redis state case: mapping shared memory of a value that was never read pulls it on first use", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 };
    f.store.set(f.keyOf(kv), v.data(), v.size());
    void* region = mmap(nullptr, faabric::util::HOST_PAGE_SIZE, PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    REQUIRE(region != MAP_FAILED);
    kv->mapSharedMemory(region, 0, 1);
    kv->get(); // makes sure the value is there
    auto* bytes = static_cast<uint8_t*>(region);
    REQUIRE(std::vector<uint8_t>(bytes, bytes + 5) == v);
    kv->unmapSharedMemory(region);
    munmap(region, faabric::util::HOST_PAGE_SIZE);
}

TEST_CASE("redis state case: pulling initialises the local copy from the store", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(6);
    REQUIRE_EQ(kv->size(), 6u);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4, 5 };
    f.store.set(f.keyOf(kv), v);
    kv->pull();
    REQUIRE_EQ(kv->size(), 6u);
    uint8_t* p = kv->get();
    REQUIRE(std::vector<uint8_t>(p, p + 6) == v);
}

TEST_CASE("redis state case: deletion removes the value from the store", "[state][redis][cases]")
{
    RedisState f;
    auto kv = f.kvOf(5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 };
    kv->set(v.data());
    kv->pushFull();
    std::string storeKey = f.keyOf(kv);
    REQUIRE(f.store.get(storeKey) == v);
    f.state.deleteKV(kv->user, kv->key);
    REQUIRE(f.store.get(storeKey).empty());
}

TEST_CASE("util state helpers: key naming and double masks", "[util][state][cases]")
{
    REQUIRE_EQ(faabric::util::keyForUser("demo", "abc"), std::string("demo_abc"));
    REQUIRE_THROWS(faabric::util::keyForUser("", "abc"));
    REQUIRE_THROWS(faabric::util::keyForUser("demo", ""));
    std::vector<unsigned int> mask(8, 0);
    faabric::util::maskDouble(mask.data(), 1);
    faabric::util::maskDouble(mask.data(), 3);
    REQUIRE(mask == (std::vector<unsigned int>{ 0, 0, STATE_MASK_32, STATE_MASK_32, 0, 0, STATE_MASK_32, STATE_MASK_32 }));
}
