// State, Redis emulation and snapshot service tests (strategy: reference
// tests/test/state/*.cpp, tests/test/redis/test_redis.cpp,
// tests/test/snapshot/*.cpp)
#include "fixtures.h"

#include <faabric/redis/Redis.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/state/InMemoryStateKeyValue.h>
#include <faabric/state/InMemoryStateRegistry.h>
#include <faabric/state/RedisStateKeyValue.h>
#include <faabric/state/State.h>
#include <faabric/state/StateClient.h>
#include <faabric/state/StateServer.h>
#include <faabric/util/bytes.h>
#include <faabric/util/memory.h>

#include <thread>

using namespace faabric::state;

TEST_CASE("redis emulation: strings, counters, ranges", "[redis]")
{
    auto& r = faabric::redis::Redis::getQueue();
    r.flushAll();
    r.ping();
    r.set("k", { 1, 2, 3, 4 });
    REQUIRE(r.get("k") == (std::vector<uint8_t>{ 1, 2, 3, 4 }));
    REQUIRE_EQ(r.strlen("k"), 4u);
    uint8_t buf[4] = { 0 };
    r.get("k", buf, 4);
    REQUIRE_EQ(buf[3], 4);
    r.del("k");
    REQUIRE(r.get("k").empty());

    REQUIRE_EQ(r.incr("ctr"), 1);
    REQUIRE_EQ(r.incrByLong("ctr", 10), 11);
    REQUIRE_EQ(r.decr("ctr"), 10);
    REQUIRE_EQ(r.decrByLong("ctr", 4), 6);
    REQUIRE_EQ(r.getCounter("ctr"), 6);
    r.setLong("lng", 123456789012L);
    REQUIRE_EQ(r.getLong("lng"), 123456789012L);

    std::vector<uint8_t> base(10, 0);
    r.set("rng", base);
    uint8_t patch[3] = { 7, 8, 9 };
    r.setRange("rng", 4, patch, 3);
    uint8_t out[5];
    r.getRange("rng", out, 5, 3, 7);
    REQUIRE_EQ(out[0], 0);
    REQUIRE_EQ(out[1], 7);
    REQUIRE_EQ(out[3], 9);
    r.setRangePipeline("rng", 0, patch, 3);
    r.flushPipeline(1);
    REQUIRE_EQ(r.get("rng")[2], 9);
}

TEST_CASE("redis emulation: sets, lists, queues, locks", "[redis]")
{
    auto& r = faabric::redis::Redis::getQueue();
    r.flushAll();
    r.sadd("s1", "a");
    r.sadd("s1", "b");
    r.sadd("s1", "b");
    r.sadd("s2", "b");
    r.sadd("s2", "c");
    REQUIRE_EQ(r.scard("s1"), 2);
    REQUIRE(r.sismember("s1", "a"));
    REQUIRE(!r.sismember("s1", "c"));
    REQUIRE(r.sdiff("s1", "s2") == (std::set<std::string>{ "a" }));
    REQUIRE(r.sinter("s1", "s2") == (std::set<std::string>{ "b" }));
    REQUIRE(r.smembers("s2").count(r.srandmember("s2")) == 1);
    r.srem("s1", "a");
    REQUIRE_EQ(r.scard("s1"), 1);

    r.enqueue("q", "first");
    r.enqueue("q", "second");
    REQUIRE_EQ(r.listLength("q"), 2);
    REQUIRE_EQ(r.dequeue("q"), std::string("first"));
    REQUIRE_EQ(r.dequeue("q"), std::string("second"));
    REQUIRE_THROWS(r.dequeue("q", 30));
    // Blocking dequeue woken by another thread
    std::thread producer([&] {
        std::this_thread::sleep_for(std::chrono::milliseconds(30));
        faabric::redis::Redis::getQueue().enqueueBytes("bq", { 5, 6 });
    });
    auto bytes = r.dequeueBytes("bq", 2000);
    producer.join();
    REQUIRE(bytes == (std::vector<uint8_t>{ 5, 6 }));
    // Longs are stored the way a redis client formats them (decimal)
    char vals[16] = { 0 };
    r.rpushLong("longs", 10);
    r.rpushLong("longs", 20);
    r.lpushLong("longs", 5);
    r.dequeueMultiple("longs", (uint8_t*)vals, sizeof(vals), 3);
    REQUIRE_EQ(std::string(vals), std::string("51020"));
    REQUIRE_EQ(r.listLength("longs"), 3);

    // Locks
    uint32_t id = r.acquireLock("lockme", 5);
    REQUIRE(id > 0);
    REQUIRE_EQ(r.acquireLock("lockme", 5), 0u);
    r.releaseLock("lockme", id + 1); // wrong owner: no-op
    REQUIRE_EQ(r.acquireLock("lockme", 5), 0u);
    r.releaseLock("lockme", id);
    uint32_t id2 = r.acquireLock("lockme", 5);
    REQUIRE(id2 > 0);
    r.releaseLock("lockme", id2);
    REQUIRE(r.setnxex("once", 1, 1));
    REQUIRE(!r.setnxex("once", 2, 1));
    r.expire("once", 0);
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
    REQUIRE(r.setnxex("once", 3, 1));
    // Queue and state roles are separate stores
    REQUIRE(faabric::redis::Redis::getState().get("once").empty());
}

namespace {
struct StateFixture
{
    State& mainState = getGlobalState();
    StateServer server;
    // A second host's view of the state
    State remoteState;

    StateFixture()
      : server(getGlobalState())
      , remoteState("otherhost")
    {
        faabric::util::getSystemConfig().reset();
        mainState.forceClearAll(true);
        getInMemoryStateRegistry().clear();
        faabric::redis::Redis::getState().flushAll();
        server.start();
    }

    ~StateFixture()
    {
        server.stop();
        mainState.forceClearAll(true);
        remoteState.forceClearAll(false);
        getInMemoryStateRegistry().clear();
        faabric::util::getSystemConfig().reset();
    }
};
}

TEST_CASE("state: in-memory main and remote replicas", "[state]")
{
    StateFixture f;
    size_t size = 3 * STATE_STREAMING_CHUNK_SIZE + 123;
    std::vector<uint8_t> values(size);
    for (size_t i = 0; i < size; i++) {
        values[i] = (uint8_t)(i * 31);
    }
    auto mainKv = f.mainState.getKV("demo", "big", size);
    mainKv->set(values.data());
    mainKv->pushFull();
    REQUIRE_EQ(f.mainState.getKVCount(), 1u);
    REQUIRE_EQ(f.mainState.getStateSize("demo", "big"), size);

    // The other host learns the size and pulls through the state server
    REQUIRE_EQ(f.remoteState.getStateSize("demo", "big"), size);
    auto remoteKv = f.remoteState.getKV("demo", "big");
    // Size-less replicas configure themselves on first use
    std::vector<uint8_t> pulled(size, 0);
    remoteKv->get(pulled.data());
    REQUIRE_EQ(remoteKv->size(), size);
    REQUIRE(pulled == values);

    // Chunked lazy pull only brings what is asked for
    auto remoteChunkKv = State("thirdhost").getKV("demo", "big");
    std::vector<uint8_t> part(100);
    remoteChunkKv->getChunk(STATE_STREAMING_CHUNK_SIZE + 10, part.data(), 100);
    REQUIRE_EQ(part[0], values[STATE_STREAMING_CHUNK_SIZE + 10]);

    // Remote partial update reaches main
    std::vector<uint8_t> patch(64, 0xee);
    remoteKv->setChunk(2 * STATE_STREAMING_CHUNK_SIZE + 5, patch.data(), patch.size());
    remoteKv->pushPartial();
    std::vector<uint8_t> check(64);
    mainKv->getChunk(2 * STATE_STREAMING_CHUNK_SIZE + 5, check.data(), 64);
    REQUIRE(check == patch);
    // ...and untouched bytes stay
    REQUIRE_EQ(*mainKv->getChunk(3, 1), values[3]);

    // Appends accumulate on main from any host
    std::vector<uint8_t> a = { 1, 1, 1 }, b = { 2, 2, 2 };
    auto mainApp = f.mainState.getKV("demo", "log", 3);
    mainApp->append(a.data(), 3);
    auto remoteApp = f.remoteState.getKV("demo", "log", 3);
    remoteApp->append(b.data(), 3);
    std::vector<uint8_t> appended(6);
    remoteApp->getAppended(appended.data(), 6, 2);
    REQUIRE(appended == (std::vector<uint8_t>{ 1, 1, 1, 2, 2, 2 }));
    remoteApp->clearAppended();
    REQUIRE_THROWS(mainApp->getAppended(appended.data(), 6, 2));

    // Deletion
    f.remoteState.deleteKV("demo", "big");
    REQUIRE_EQ(f.mainState.getKVCount(), 1u);
}

TEST_CASE("state: shared-memory mapping and locks", "[state]")
{
    StateFixture f;
    size_t size = 2 * faabric::util::HOST_PAGE_SIZE;
    auto kv = f.mainState.getKV("demo", "mapped", size);
    std::vector<uint8_t> init(size, 3);
    kv->set(init.data());
    auto region = faabric::util::allocatePrivateMemory(size);
    kv->mapSharedMemory(region.get(), 0, 2);
    REQUIRE_EQ(region[10], 3);
    region[10] = 9; // writes through to the KV
    REQUIRE_EQ(*kv->getChunk(10, 1), 9);
    // the KV unmapped the range: the region must not unmap it again later
    // (by then the addresses may belong to somebody else)
    kv->unmapSharedMemory(region.release());

    // Write lock excludes readers
    kv->lockWrite();
    std::atomic<bool> got{ false };
    std::thread reader([&] {
        kv->lockRead();
        got = true;
        kv->unlockRead();
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    REQUIRE(!got.load());
    kv->unlockWrite();
    reader.join();
    REQUIRE(got.load());
    // A size-less KV nobody has created fails on first use
    auto ghost = f.mainState.getKV("demo", "nosize");
    REQUIRE_THROWS(ghost->get());
}

TEST_CASE("state: chunk bounds, dirty flags, masks and lazy pulls", "[state]")
{
    StateFixture f;
    const size_t page = faabric::util::HOST_PAGE_SIZE;
    // Value smaller than its page-rounded backing store
    const size_t size = page + 100;
    auto mainKv = f.mainState.getKV("demo", "bounds", size);
    REQUIRE_EQ(mainKv->size(), size);
    REQUIRE_EQ(mainKv->getSharedMemorySize(), 2 * page);
    std::vector<uint8_t> values(size, 5);
    mainKv->set(values.data());
    // A chunk may run past the value, up to the end of the allocation...
    std::vector<uint8_t> tail(50, 7);
    mainKv->setChunk((long)(2 * page - 50), tail.data(), 50);
    // ...but not past it
    REQUIRE_THROWS(mainKv->setChunk((long)(2 * page - 49), tail.data(), 50));
    REQUIRE_THROWS(mainKv->getChunk((long)(2 * page), 1));

    // all chunks tile the value in streaming-size pieces
    const size_t big = 2 * STATE_STREAMING_CHUNK_SIZE + 10;
    auto bigKv = f.mainState.getKV("demo", "tiles", big);
    std::vector<uint8_t> bigValues(big);
    for (size_t i = 0; i < big; i++) {
        bigValues[i] = (uint8_t)(i % 251);
    }
    bigKv->set(bigValues.data());
    auto chunks = bigKv->getAllChunks();
    REQUIRE_EQ(chunks.size(), (size_t)3);
    REQUIRE_EQ(chunks[0].offset, 0L);
    REQUIRE_EQ(chunks[1].offset, (long)STATE_STREAMING_CHUNK_SIZE);
    REQUIRE_EQ(chunks[2].length, (size_t)10);
    REQUIRE_EQ(chunks[2].data[0], bigValues[2 * STATE_STREAMING_CHUNK_SIZE]);

    // Replica on another host: lazy get pulls once, pull() always re-pulls
    State other("hostX");
    auto replica = other.getKV("demo", "tiles", big);
    std::vector<uint8_t> got(big, 0);
    replica->get(got.data());
    REQUIRE(got == bigValues);
    uint8_t nine = 9;
    bigKv->setChunk(5, &nine, 1);
    replica->get(got.data());
    REQUIRE_EQ(got[5], bigValues[5]); // still the first pull
    replica->pull();
    replica->get(got.data());
    REQUIRE_EQ(got[5], 9);

    // A fresh replica mapped into memory is NOT pulled by the mapping
    auto mapped = State("hostY").getKV("demo", "tiles", big);
    auto region = faabric::util::allocatePrivateMemory(faabric::util::getRequiredHostPages(big) * faabric::util::HOST_PAGE_SIZE);
    mapped->mapSharedMemory(region.get(), 0, (long)faabric::util::getRequiredHostPages(big));
    REQUIRE_EQ(region[100], 0);
    mapped->unmapSharedMemory(region.release());

    // Pushes are no-ops unless something is dirty
    auto quiet = State("hostZ").getKV("demo", "tiles", big);
    quiet->get(got.data());
    quiet->pushFull();
    quiet->pushPartial();
    REQUIRE_EQ(*bigKv->getChunk(5, 1), 9);
    // Writing through a mapped pointer needs an explicit flag
    uint8_t* raw = quiet->get();
    raw[10] = 77;
    raw[20] = 78;
    quiet->pushPartial();
    REQUIRE_EQ(*bigKv->getChunk(10, 1), bigValues[10]);
    quiet->flagChunkDirty(10, 1);
    quiet->pushPartial();
    REQUIRE_EQ(*bigKv->getChunk(10, 1), 77);
    REQUIRE_EQ(*bigKv->getChunk(20, 1), bigValues[20]);
    quiet->flagDirty();
    quiet->pushFull();
    REQUIRE_EQ(*bigKv->getChunk(20, 1), 78);

    // Partial push driven by a mask held in another KV
    auto mask = State("hostZ").getKV("demo", "tiles-mask", big);
    std::vector<uint8_t> maskBytes(big, 0);
    std::fill(maskBytes.begin() + 1000, maskBytes.begin() + 1010, 1);
    mask->set(maskBytes.data());
    raw[1005] = 55;
    raw[2000] = 56; // outside the mask: stays local
    quiet->pushPartialMask(mask); // clean value: nothing goes out
    REQUIRE_EQ(*bigKv->getChunk(1005, 1), bigValues[1005]);
    quiet->flagDirty();
    quiet->pushPartialMask(mask);
    REQUIRE_EQ(*bigKv->getChunk(1005, 1), 55);
    REQUIRE_EQ(*bigKv->getChunk(2000, 1), bigValues[2000]);
    auto wrongMask = State("hostZ").getKV("demo", "small-mask", 10);
    REQUIRE_THROWS(quiet->pushPartialMask(wrongMask));

    // Local deletion forgets the replica, not the value
    size_t before = f.mainState.getKVCount();
    f.mainState.deleteKVLocally("demo", "bounds");
    REQUIRE_EQ(f.mainState.getKVCount(), before - 1);
    REQUIRE_EQ(f.mainState.getThisIP(), faabric::util::getSystemConfig().endpointHost);
}

TEST_CASE("state: redis-backed mode", "[state][redis]")
{
    StateFixture f;
    faabric::util::getSystemConfig().stateMode = "redis";
    State a("hostA"), b("hostB");
    std::vector<uint8_t> v = { 9, 8, 7, 6, 5 };
    auto kvA = a.getKV("demo", "r", v.size());
    kvA->set(v.data());
    kvA->pushFull();
    REQUIRE_EQ(b.getStateSize("demo", "r"), v.size());
    auto kvB = b.getKV("demo", "r", v.size());
    std::vector<uint8_t> got(v.size());
    kvB->get(got.data());
    REQUIRE(got == v);
    uint8_t patch[2] = { 1, 2 };
    kvB->setChunk(1, patch, 2);
    kvB->pushPartial();
    kvA->pull();
    kvA->get(got.data());
    REQUIRE(got == (std::vector<uint8_t>{ 9, 1, 2, 6, 5 }));
    a.deleteKV("demo", "r");
    REQUIRE_EQ(State("hostC").getStateSize("demo", "r"), 0u);
}

TEST_CASE("snapshots: push, update, thread results, delete over RPC", "[snapshot]")
{
    tests::ClusterFixture f(2);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    size_t size = 4 * faabric::util::HOST_PAGE_SIZE;
    auto snap = std::make_shared<faabric::util::SnapshotData>(size);
    std::vector<uint8_t> content(size, 1);
    snap->copyInData(content);
    snap->addMergeRegion(64, sizeof(int), faabric::util::SnapshotDataType::Int, faabric::util::SnapshotMergeOperation::Sum);

    // The client talks to this very host: the server registers a copy
    faabric::snapshot::SnapshotClient cli(f.conf.endpointHost);
    cli.pushSnapshot("snapA", snap);
    REQUIRE(reg.snapshotExists("snapA"));
    auto received = reg.getSnapshot("snapA");
    REQUIRE(received.get() != snap.get());
    REQUIRE_EQ(received->getSize(), size);
    REQUIRE_EQ(received->getMergeRegions().size(), 1u);
    REQUIRE_EQ(*received->getDataPtr(100), 1);
    REQUIRE_EQ(received->getTrackedChanges().size(), 0u);

    // Updates are applied immediately and replace the merge regions
    std::vector<uint8_t> bytes(16, 0x7f);
    int delta = 5;
    std::vector<faabric::util::SnapshotDiff> diffs;
    diffs.emplace_back(faabric::util::SnapshotDataType::Raw, faabric::util::SnapshotMergeOperation::Bytewise, 1000, bytes);
    diffs.emplace_back(faabric::util::SnapshotDataType::Int,
                       faabric::util::SnapshotMergeOperation::Sum,
                       64,
                       std::span<const uint8_t>((const uint8_t*)&delta, sizeof(int)));
    snap->clearMergeRegions();
    cli.pushSnapshotUpdate("snapA", snap, diffs);
    REQUIRE_EQ(*received->getDataPtr(1000), 0x7f);
    REQUIRE_EQ(faabric::util::unalignedRead<int>(received->getDataPtr(64)), 0x01010101 + 5);
    REQUIRE_EQ(received->getMergeRegions().size(), 0u);

    // Thread results queue their diffs; the return value reaches the waiter
    // through the planner
    auto threadReq = faabric::util::batchExecFactory("demo", "thr", 1);
    uint32_t msgId = threadReq->messages(0).id();
    std::vector<faabric::util::SnapshotDiff> threadDiffs;
    threadDiffs.emplace_back(faabric::util::SnapshotDataType::Raw, faabric::util::SnapshotMergeOperation::Bytewise, 2000, bytes);
    cli.pushThreadResult(threadReq->appid(), msgId, 42, "snapA", threadDiffs);
    REQUIRE_EQ(received->getQueuedDiffsCount(), 1u);
    {
        faabric::HostResources res;
        res.set_slots(2);
        res.set_usedslots(1);
        f.sch.setThisHostResources(res);
        auto viaPlanner = std::make_shared<faabric::Message>(threadReq->messages(0));
        viaPlanner->set_returnvalue(42);
        viaPlanner->set_executedhost(f.conf.endpointHost);
        f.plannerCli.setMessageResult(viaPlanner);
    }
    auto results = f.sch.awaitThreadResults(threadReq, 2000);
    REQUIRE_EQ(results.size(), 1u);
    REQUIRE_EQ(results[0].first, msgId);
    REQUIRE_EQ(results[0].second, 42);
    REQUIRE_EQ(received->writeQueuedDiffs(), 1);
    REQUIRE_EQ(*received->getDataPtr(2000), 0x7f);

    cli.deleteSnapshot("snapA");
    for (int i = 0; i < 200 && reg.snapshotExists("snapA"); i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    REQUIRE(!reg.snapshotExists("snapA"));
    REQUIRE_THROWS(reg.getSnapshot("snapA"));
    REQUIRE_THROWS(reg.getSnapshot(""));

    // Mock mode records instead of sending
    faabric::util::setMockMode(true);
    faabric::snapshot::clearMockSnapshotRequests();
    faabric::snapshot::SnapshotClient mockCli("elsewhere");
    mockCli.pushSnapshot("snapB", snap);
    mockCli.pushSnapshotUpdate("snapB", snap, diffs);
    mockCli.deleteSnapshot("snapB");
    REQUIRE_EQ(faabric::snapshot::getSnapshotPushes().size(), 1u);
    REQUIRE_EQ(faabric::snapshot::getSnapshotDiffPushes().size(), 1u);
    REQUIRE_EQ(faabric::snapshot::getSnapshotDiffPushes()[0].second->diffs.size(), 2u);
    REQUIRE_EQ(faabric::snapshot::getSnapshotDeletes().size(), 1u);
    faabric::snapshot::clearMockSnapshotRequests();
    faabric::util::setMockMode(false);
}

// ---------------------------------------------------------------------------
// Checkpoint files (no reference counterpart: its snapshots never leave memory,
// SURVEY §5.4)
// ---------------------------------------------------------------------------
#include <filesystem>

TEST_CASE("snapshots: checkpoint files round-trip image and merge regions", "[snapshot][checkpoint]")
{
    using namespace faabric::util;
    const std::string dir = "/tmp/fb_ckpt_" + std::to_string(getpid());
    std::filesystem::remove_all(dir);
    std::filesystem::create_directories(dir);

    const size_t size = 5 * HOST_PAGE_SIZE + 123;
    std::vector<uint8_t> bytes(size);
    for (size_t i = 0; i < size; i++) {
        bytes[i] = (uint8_t)(i * 7 + 3);
    }
    SnapshotData snap(bytes, 16 * HOST_PAGE_SIZE);
    snap.addMergeRegion(64, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    snap.addMergeRegion(4096, 0, SnapshotDataType::Raw, SnapshotMergeOperation::XOR);
    const std::string path = dir + "/one.snap";
    snap.writeToFile(path);
    // no temp file left behind
    size_t nFiles = 0;
    for (auto& e : std::filesystem::directory_iterator(dir)) {
        (void)e;
        nFiles++;
    }
    REQUIRE_EQ(nFiles, (size_t)1);

    auto back = SnapshotData::readFromFile(path);
    REQUIRE_EQ(back->getSize(), size);
    REQUIRE_EQ(back->getMaxSize(), (size_t)(16 * HOST_PAGE_SIZE));
    REQUIRE(back->getDataCopy() == bytes);
    auto regions = back->getMergeRegions();
    REQUIRE_EQ(regions.size(), (size_t)2);
    REQUIRE_EQ(regions[0].offset, (uint64_t)64);
    REQUIRE(regions[0].dataType == SnapshotDataType::Int);
    REQUIRE(regions[0].operation == SnapshotMergeOperation::Sum);
    REQUIRE_EQ(regions[1].length, (uint64_t)0);
    REQUIRE(regions[1].operation == SnapshotMergeOperation::XOR);
    // the restored image is a full snapshot: it can grow and be mapped
    std::vector<uint8_t> extra(HOST_PAGE_SIZE, 9);
    back->copyInData(extra, 8 * HOST_PAGE_SIZE);
    REQUIRE_EQ(back->getSize(), (size_t)(9 * HOST_PAGE_SIZE));
    MemoryRegion mem = allocatePrivateMemory(back->getSize());
    back->mapToMemory({ mem.get(), back->getSize() });
    REQUIRE(std::equal(bytes.begin(), bytes.end(), mem.get()));

    // overwrite in place, empty snapshots, error cases
    SnapshotData empty;
    empty.writeToFile(path);
    REQUIRE_EQ(SnapshotData::readFromFile(path)->getSize(), (size_t)0);
    REQUIRE_THROWS(SnapshotData::readFromFile(dir + "/missing.snap"));
    writeBytesToFile(dir + "/junk.snap", std::vector<uint8_t>(100, 1));
    REQUIRE_THROWS(SnapshotData::readFromFile(dir + "/junk.snap"));
    snap.writeToFile(path);
    std::filesystem::resize_file(path, std::filesystem::file_size(path) - 10);
    REQUIRE_THROWS(SnapshotData::readFromFile(path));
    REQUIRE_THROWS(snap.writeToFile(dir + "/no/such/dir/x.snap"));
    std::filesystem::remove_all(dir);
}

TEST_CASE("snapshots: registry checkpoints to a directory and restores", "[snapshot][checkpoint]")
{
    using namespace faabric::util;
    const std::string dir = "/tmp/fb_ckpt_reg_" + std::to_string(getpid());
    std::filesystem::remove_all(dir);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    reg.clear();
    std::vector<uint8_t> a(3000, 1), b(HOST_PAGE_SIZE * 2, 2);
    auto sa = std::make_shared<SnapshotData>(a);
    auto sb = std::make_shared<SnapshotData>(b);
    sb->addMergeRegion(0, 8, SnapshotDataType::Long, SnapshotMergeOperation::Max);
    // keys with characters that are not file-name safe
    reg.registerSnapshot("demo/echo_123", sa);
    reg.registerSnapshot("migration_77", sb);
    REQUIRE_EQ(reg.checkpointToDir(dir), (size_t)2);
    // unrelated files are ignored on restore
    writeBytesToFile(dir + "/README", stringToBytes("not a snapshot"));
    writeBytesToFile(dir + "/zz.snap", stringToBytes("name is not hex"));

    reg.clear();
    REQUIRE_EQ(reg.getSnapshotCount(), (size_t)0);
    REQUIRE_EQ(reg.restoreFromDir(dir), (size_t)2);
    REQUIRE(reg.snapshotExists("demo/echo_123"));
    REQUIRE(reg.getSnapshot("demo/echo_123")->getDataCopy() == a);
    auto rb = reg.getSnapshot("migration_77");
    REQUIRE(rb->getDataCopy() == b);
    REQUIRE_EQ(rb->getMergeRegions().size(), (size_t)1);
    REQUIRE(rb->getMergeRegions()[0].operation == SnapshotMergeOperation::Max);
    REQUIRE_EQ(reg.restoreFromDir(dir + "/nothing-here"), (size_t)0);
    // a later checkpoint drops the files of deleted snapshots
    reg.deleteSnapshot("migration_77");
    REQUIRE_EQ(reg.checkpointToDir(dir), (size_t)1);
    reg.clear();
    REQUIRE_EQ(reg.restoreFromDir(dir), (size_t)1);
    REQUIRE(!reg.snapshotExists("migration_77"));
    reg.clear();
    std::filesystem::remove_all(dir);
}
