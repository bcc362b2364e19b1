// Util layer tests (strategy mirrors reference tests/test/util/*.cpp)
#include "harness.h"

#include <faabric/util/barrier.h>
#include <faabric/util/batch.h>
#include <faabric/util/bytes.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/delta.h>
#include <faabric/util/dirty.h>
#include <faabric/util/environment.h>
#include <faabric/util/func.h>
#include <faabric/util/gids.h>
#include <faabric/util/json.h>
#include <faabric/util/latch.h>
#include <faabric/util/locks.h>
#include <faabric/util/memory.h>
#include <faabric/util/queue.h>
#include <faabric/util/random.h>
#include <faabric/util/snapshot.h>
#include <faabric/util/string_tools.h>

#include <atomic>
#include <set>
#include <thread>

#ifndef SLEEP_MS
#define SLEEP_MS(ms) std::this_thread::sleep_for(std::chrono::milliseconds(ms))
#endif

using namespace faabric::util;

TEST_CASE("bytes: string and hex round trips", "[util]")
{
    std::string s = "hello \x01 world";
    REQUIRE_EQ(bytesToString(stringToBytes(s)), s);
    std::vector<uint8_t> b = { 0x00, 0xab, 0x10, 0xff };
    REQUIRE_EQ(byteArrayToHexString(b.data(), (int)b.size()), std::string("00ab10ff"));
    REQUIRE(hexStringToByteArray("00ab10ff") == b);
    std::vector<uint8_t> z = { 1, 2, 0, 3, 0, 0 };
    trimTrailingZeros(z);
    REQUIRE_EQ(z.size(), 4u);
    uint8_t buf[3];
    REQUIRE_EQ(safeCopyToBuffer(b, buf, 3), 3);
    REQUIRE_EQ(buf[1], 0xab);
    REQUIRE_EQ(unalignedRead<uint32_t>(valueToBytes<uint32_t>(0xdeadbeef).data()), 0xdeadbeefu);
    REQUIRE_EQ(formatByteArrayToIntString({ 1, 2, 3 }), std::string("[1, 2, 3]"));
}

TEST_CASE("string tools", "[util]")
{
    REQUIRE(isAllWhitespace("  \t\n"));
    REQUIRE(!isAllWhitespace(" a "));
    REQUIRE(startsWith("foobar", "foo"));
    REQUIRE(!startsWith("foobar", "bar"));
    REQUIRE(endsWith("foobar", "bar"));
    REQUIRE(contains("foobar", "oba"));
    REQUIRE_EQ(removeSubstr("foobar", "ob"), std::string("foar"));
    REQUIRE(stringIsInt("1234"));
    REQUIRE(!stringIsInt("12a"));
    auto parts = splitString("a,b,,c", ',');
    REQUIRE_EQ(parts.size(), 3u);
    REQUIRE_EQ(parts[2], std::string("c"));
    REQUIRE_EQ(trim("  x y \n"), std::string("x y"));
    REQUIRE_EQ(toLower("AbC"), std::string("abc"));
}

TEST_CASE("gids are unique across threads", "[util]")
{
    std::vector<std::thread> ts;
    std::mutex mx;
    std::set<unsigned int> all;
    for (int t = 0; t < 8; t++) {
        ts.emplace_back([&] {
            std::vector<unsigned int> mine;
            for (int i = 0; i < 1000; i++) {
                mine.push_back(generateGid());
            }
            std::lock_guard<std::mutex> lk(mx);
            all.insert(mine.begin(), mine.end());
        });
    }
    for (auto& t : ts) {
        t.join();
    }
    REQUIRE_EQ(all.size(), 8000u);
    REQUIRE(all.count(0) == 0);
}

TEST_CASE("config reads environment and resets", "[util]")
{
    auto& conf = getSystemConfig();
    conf.reset();
    REQUIRE_EQ(conf.batchSchedulerMode, std::string("bin-pack"));
    REQUIRE_EQ(conf.stateMode, std::string("inmemory"));
    setEnvVar("BATCH_SCHEDULER_MODE", "compact");
    setEnvVar("BOUND_TIMEOUT", "1234");
    conf.reset();
    REQUIRE_EQ(conf.batchSchedulerMode, std::string("compact"));
    REQUIRE_EQ(conf.boundTimeout, 1234);
    unsetEnvVar("BATCH_SCHEDULER_MODE");
    unsetEnvVar("BOUND_TIMEOUT");
    conf.reset();
    REQUIRE_EQ(conf.batchSchedulerMode, std::string("bin-pack"));
}

TEST_CASE("queue: fifo, timeout, drain", "[util]")
{
    Queue<int> q;
    for (int i = 0; i < 5; i++) {
        q.enqueue(i);
    }
    REQUIRE_EQ(q.size(), 5);
    for (int i = 0; i < 5; i++) {
        REQUIRE_EQ(q.dequeue(), i);
    }
    REQUIRE_THROWS(q.dequeue(20));
    std::thread t([&] {
        for (int i = 0; i < 100; i++) {
            q.enqueue(i);
        }
    });
    long sum = 0;
    for (int i = 0; i < 100; i++) {
        sum += q.dequeue(2000);
    }
    t.join();
    REQUIRE_EQ(sum, 4950);
    q.waitToDrain(100);
}

TEST_CASE("fixed capacity + spin queues across threads", "[util]")
{
    FixedCapacityQueue<int> fq(4);
    SpinLockQueue<int> sq(4);
    std::thread prod([&] {
        for (int i = 0; i < 2000; i++) {
            fq.enqueue(i);
            sq.enqueue(i * 2);
        }
    });
    for (int i = 0; i < 2000; i++) {
        REQUIRE_EQ(fq.dequeue(), i);
        REQUIRE_EQ(sq.dequeue(), i * 2);
    }
    prod.join();
    REQUIRE_THROWS(fq.dequeue(10));
    // Bounded: filling past capacity times out
    for (int i = 0; i < 4; i++) {
        fq.enqueue(i);
    }
    REQUIRE_THROWS(fq.enqueue(5, 20));
}

TEST_CASE("latch, barrier and flag waiter", "[util]")
{
    auto latch = Latch::create(4);
    auto barrier = Barrier::create(4);
    std::atomic<int> passed{ 0 };
    std::vector<std::thread> ts;
    for (int i = 0; i < 3; i++) {
        ts.emplace_back([&] {
            latch->wait();
            for (int r = 0; r < 10; r++) {
                barrier->wait();
            }
            passed++;
        });
    }
    latch->wait();
    for (int r = 0; r < 10; r++) {
        barrier->wait();
    }
    for (auto& t : ts) {
        t.join();
    }
    REQUIRE_EQ(passed.load(), 3);
    // Latch with missing participants times out
    Latch shortLatch(2, 30);
    REQUIRE_THROWS(shortLatch.wait());

    auto fw = std::make_shared<FlagWaiter>(2000);
    std::thread setter([&] {
        SLEEP_MS(20);
        fw->setFlag(true);
    });
    fw->waitOnFlag();
    setter.join();
    FlagWaiter never(30);
    REQUIRE_THROWS(never.waitOnFlag());
}

TEST_CASE("message factory and ids", "[util]")
{
    auto msg = messageFactory("demo", "echo");
    REQUIRE(msg.id() > 0);
    REQUIRE(msg.appid() > 0);
    REQUIRE_EQ(msg.user(), std::string("demo"));
    REQUIRE(!msg.mainhost().empty());
    REQUIRE_EQ(funcToString(msg, false), std::string("demo/echo"));
    REQUIRE(funcToString(msg, true).find(std::to_string(msg.id())) != std::string::npos);
    msg.set_cmdline("a b c");
    auto argv = getArgvForMessage(msg);
    REQUIRE_EQ(argv.size(), 4u);
    REQUIRE_EQ(argv[0], std::string("function.wasm"));

    auto ber = batchExecFactory("demo", "echo", 4);
    REQUIRE(isBatchExecRequestValid(ber));
    REQUIRE_EQ(ber->messages_size(), 4);
    for (int i = 0; i < 4; i++) {
        REQUIRE_EQ(ber->messages(i).appid(), ber->appid());
        REQUIRE_EQ(ber->messages(i).appidx(), i);
    }
    // Breaking an app id breaks validity
    ber->mutable_messages(2)->set_appid(1337);
    REQUIRE(!isBatchExecRequestValid(ber));
    REQUIRE(!isBatchExecRequestValid(nullptr));
    updateBatchExecAppId(ber, 99);
    REQUIRE(isBatchExecRequestValid(ber));
    REQUIRE_EQ(ber->messages(3).appid(), 99);
}

TEST_CASE("message json round trip", "[util]")
{
    auto msg = messageFactory("demo", "echo");
    msg.set_inputdata("in\x01put");
    msg.set_ismpi(true);
    msg.set_mpiworldsize(8);
    msg.set_returnvalue(-3);
    msg.add_chainedmsgids(11);
    msg.add_chainedmsgids(12);
    msg.set_recordexecgraph(true);
    (*msg.mutable_execgraphdetails())["k"] = "v";
    (*msg.mutable_intexecgraphdetails())["n"] = 7;
    std::string js = messageToJson(msg);
    faabric::Message back;
    jsonToMessage(js, &back);
    REQUIRE_EQ(back.id(), msg.id());
    REQUIRE_EQ(back.inputdata(), msg.inputdata());
    REQUIRE_EQ(back.mpiworldsize(), 8);
    REQUIRE_EQ(back.returnvalue(), -3);
    REQUIRE_EQ(back.chainedmsgids_size(), 2);
    REQUIRE_EQ(back.execgraphdetails().at("k"), std::string("v"));
    REQUIRE_EQ(back.intexecgraphdetails().at("n"), 7);
    // Wire round trip too
    faabric::Message wire;
    REQUIRE(wire.ParseFromString(msg.SerializeAsString()));
    REQUIRE_EQ(messageToJson(wire), js);
    REQUIRE_THROWS(jsonToMessage("{not json", &back));
}

TEST_CASE("page maths and dirty page merge", "[util]")
{
    REQUIRE_EQ(getRequiredHostPages(1), 1u);
    REQUIRE_EQ(getRequiredHostPages((size_t)HOST_PAGE_SIZE), 1u);
    REQUIRE_EQ(getRequiredHostPages((size_t)HOST_PAGE_SIZE + 1), 2u);
    REQUIRE_EQ(getRequiredHostPagesRoundDown((size_t)HOST_PAGE_SIZE * 2 - 1), 1u);
    REQUIRE_EQ(alignOffsetDown((size_t)HOST_PAGE_SIZE + 5), (size_t)HOST_PAGE_SIZE);
    auto c = getPageAlignedChunk(HOST_PAGE_SIZE + 10, HOST_PAGE_SIZE);
    REQUIRE_EQ(c.nBytesOffset, HOST_PAGE_SIZE);
    REQUIRE_EQ(c.nPagesLength, 2);
    REQUIRE_EQ(c.offsetRemainder, 10);
    std::vector<char> a = { 0, 1, 0 };
    mergeDirtyPages(a, { 1, 0, 0, 1 });
    REQUIRE(a == (std::vector<char>{ 1, 1, 0, 1 }));
    auto mem = allocatePrivateMemory(3 * HOST_PAGE_SIZE);
    REQUIRE(isPageAligned(mem.get()));
    REQUIRE_EQ(mem[100], 0);
}

static void checkTracker(const std::string& mode)
{
    auto& conf = getSystemConfig();
    conf.dirtyTrackingMode = mode;
    resetDirtyTracker();
    auto tracker = getDirtyTracker();
    REQUIRE_EQ(tracker->getType(), mode);
    size_t nPages = 8;
    auto mem = allocatePrivateMemory(nPages * HOST_PAGE_SIZE);
    std::span<uint8_t> region(mem.get(), nPages * HOST_PAGE_SIZE);
    // Touch everything before tracking so pages are mapped
    memset(mem.get(), 1, region.size());
    tracker->clearAll();
    tracker->startTracking(region);
    tracker->startThreadLocalTracking(region);
    mem[HOST_PAGE_SIZE * 1 + 5] = 9;
    mem[HOST_PAGE_SIZE * 6] = 9;
    // Each thread reports the pages it dirtied itself
    std::vector<char> otherPages;
    std::thread other([&] {
        tracker->startThreadLocalTracking(region);
        mem[HOST_PAGE_SIZE * 3 + 1] = 7;
        tracker->stopThreadLocalTracking(region);
        otherPages = tracker->getThreadLocalDirtyPages(region);
    });
    other.join();
    tracker->stopThreadLocalTracking(region);
    tracker->stopTracking(region);
    auto both = tracker->getBothDirtyPages(region);
    mergeDirtyPages(both, otherPages);
    REQUIRE_EQ(both.size(), nPages);
    if (mode == "none") {
        for (char p : both) {
            REQUIRE_EQ((int)p, 1);
        }
    } else {
        std::vector<char> expected(nPages, 0);
        expected[1] = expected[3] = expected[6] = 1;
        REQUIRE(both == expected);
    }
    // Memory is writable again afterwards
    mem[0] = 3;
    conf.reset();
    resetDirtyTracker();
}

TEST_CASE("dirty tracking: none", "[util][dirty]")
{
    checkTracker("none");
}

TEST_CASE("dirty tracking: segfault", "[util][dirty]")
{
    checkTracker("segfault");
}

TEST_CASE("dirty tracking: softpte", "[util][dirty]")
{
    if (!SoftPTEDirtyTracker::isSupported()) {
        SKIP_TEST("soft-dirty PTEs not available");
    }
    checkTracker("softpte");
}

TEST_CASE("dirty tracking: uffd", "[util][dirty]")
{
    if (!UffdDirtyTracker::isSupported()) {
        SKIP_TEST("userfaultfd write-protect not available");
    }
    checkTracker("uffd");
}

// Every userfaultfd flavour the reference names (write-protect vs missing-page
// faults, SIGBUS vs event thread) must report the same pages
TEST_CASE("dirty tracking: uffd-wp, uffd-thread, uffd-thread-wp", "[util][dirty]")
{
    if (!UffdDirtyTracker::isSupported()) {
        SKIP_TEST("userfaultfd write-protect not available");
    }
    for (const char* mode : { "uffd-wp", "uffd-thread", "uffd-thread-wp" }) {
        checkTracker(mode);
    }
}

TEST_CASE("dirty tracking: unknown mode is rejected", "[util][dirty]")
{
    auto& conf = getSystemConfig();
    conf.dirtyTrackingMode = "telepathy";
    // (the tracker is rebuilt eagerly)
    REQUIRE_THROWS(resetDirtyTracker());
    conf.reset();
    resetDirtyTracker();
    REQUIRE(getDirtyTracker() != nullptr);
}

TEST_CASE("snapshot: typed merge diffs and application", "[util][snapshot]")
{
    getSystemConfig().diffingMode = "bytewise";
    size_t size = 4 * HOST_PAGE_SIZE;
    auto snap = std::make_shared<SnapshotData>(size);
    std::vector<uint8_t> init(size, 0);
    int base = 100;
    memcpy(init.data() + 64, &base, sizeof(int));
    double dbase = 2.5;
    memcpy(init.data() + HOST_PAGE_SIZE + 8, &dbase, sizeof(double));
    snap->copyInData(init);
    snap->clearTrackedChanges();

    auto mem = allocatePrivateMemory(size);
    std::span<uint8_t> memView(mem.get(), size);
    snap->mapToMemory(memView);
    REQUIRE_EQ(unalignedRead<int>(mem.get() + 64), 100);

    snap->addMergeRegion(64, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    snap->addMergeRegion(HOST_PAGE_SIZE + 8, sizeof(double), SnapshotDataType::Double, SnapshotMergeOperation::Product);
    snap->addMergeRegion(2 * HOST_PAGE_SIZE, 16, SnapshotDataType::Raw, SnapshotMergeOperation::Ignore);
    snap->fillGapsWithBytewiseRegions();

    // Worker-side modifications
    unalignedWrite<int>(130, mem.get() + 64);
    unalignedWrite<double>(10.0, mem.get() + HOST_PAGE_SIZE + 8);
    mem[2 * HOST_PAGE_SIZE + 3] = 0xff;         // ignored
    mem[3 * HOST_PAGE_SIZE + 100] = 0x42;       // bytewise
    mem[3 * HOST_PAGE_SIZE + 101] = 0x43;
    std::vector<char> dirty(4, 1);
    auto diffs = snap->diffWithDirtyRegions(memView, dirty);
    REQUIRE_EQ(diffs.size(), 3u);
    std::map<uint64_t, SnapshotDiff> byOff;
    for (auto& d : diffs) {
        byOff[d.getOffset()] = d;
    }
    REQUIRE(byOff.count(64));
    REQUIRE_EQ((int)byOff[64].getOperation(), (int)SnapshotMergeOperation::Sum);
    REQUIRE_EQ(unalignedRead<int>(byOff[64].getData().data()), 30);
    REQUIRE(byOff.count(HOST_PAGE_SIZE + 8));
    REQUIRE_NEAR(unalignedRead<double>(byOff[HOST_PAGE_SIZE + 8].getData().data()), 4.0, 1e-12);
    REQUIRE(byOff.count(3 * HOST_PAGE_SIZE + 100));
    REQUIRE_EQ(byOff[3 * HOST_PAGE_SIZE + 100].getData().size(), 2u);

    // Two workers' worth of diffs applied on the main copy
    snap->queueDiffs(diffs);
    snap->queueDiffs(diffs);
    REQUIRE_EQ(snap->getQueuedDiffsCount(), 6u);
    REQUIRE_EQ(snap->writeQueuedDiffs(), 6);
    REQUIRE_EQ(unalignedRead<int>(snap->getDataPtr(64)), 160);
    REQUIRE_NEAR(unalignedRead<double>(snap->getDataPtr(HOST_PAGE_SIZE + 8)), 40.0, 1e-9);
    REQUIRE_EQ(*snap->getDataPtr(2 * HOST_PAGE_SIZE + 3), 0);
    REQUIRE_EQ(*snap->getDataPtr(3 * HOST_PAGE_SIZE + 100), 0x42);
    // Tracked changes cover what was written
    auto tracked = snap->getTrackedChanges();
    REQUIRE(tracked.size() >= 3);
    getSystemConfig().reset();
}

TEST_CASE("snapshot: xor gap filling diffs whole dirty pages", "[util][snapshot]")
{
    REQUIRE_EQ(getSystemConfig().diffingMode, std::string("xor"));
    size_t size = 4 * HOST_PAGE_SIZE;
    auto snap = std::make_shared<SnapshotData>(size);
    std::vector<uint8_t> init(size, 0x11);
    snap->copyInData(init);
    auto mem = allocatePrivateMemory(size);
    std::span<uint8_t> memView(mem.get(), size);
    snap->mapToMemory(memView);
    snap->fillGapsWithBytewiseRegions();
    mem[HOST_PAGE_SIZE + 9] = 0x33;
    auto diffs = snap->diffWithDirtyRegions(memView, { 0, 1, 0, 0 });
    REQUIRE_EQ(diffs.size(), 1u);
    REQUIRE_EQ((int)diffs[0].getOperation(), (int)SnapshotMergeOperation::XOR);
    REQUIRE_EQ(diffs[0].getOffset(), (uint64_t)HOST_PAGE_SIZE);
    REQUIRE_EQ(diffs[0].getData().size(), (size_t)HOST_PAGE_SIZE);
    REQUIRE_EQ(diffs[0].getData()[9], 0x22);
    REQUIRE_EQ(diffs[0].getData()[10], 0);
    snap->applyDiffs(diffs);
    REQUIRE_EQ(*snap->getDataPtr(HOST_PAGE_SIZE + 9), 0x33);
    REQUIRE_EQ(*snap->getDataPtr(HOST_PAGE_SIZE + 10), 0x11);
}

TEST_CASE("snapshot: xor mode, growth and bounds", "[util][snapshot]")
{
    size_t size = 2 * HOST_PAGE_SIZE;
    size_t maxSize = 8 * HOST_PAGE_SIZE;
    SnapshotData snap(size, maxSize);
    REQUIRE_EQ(snap.getSize(), size);
    REQUIRE_EQ(snap.getMaxSize(), maxSize);
    std::vector<uint8_t> a(64, 0xf0), b(64, 0x0f);
    snap.copyInData(a, 128);
    snap.applyDiff(SnapshotDiff(SnapshotDataType::Raw, SnapshotMergeOperation::XOR, 128, b));
    REQUIRE_EQ(*snap.getDataPtr(128), 0xff);
    // Writing beyond the size extends it, beyond the max throws
    snap.copyInData(a, 3 * HOST_PAGE_SIZE);
    REQUIRE_EQ(snap.getSize(), 3 * HOST_PAGE_SIZE + 64);
    REQUIRE_THROWS(snap.copyInData(a, maxSize));
    REQUIRE_THROWS(snap.getDataCopy(maxSize, 10));

    // Diffs beyond original size (memory grew on the worker) are raw overwrites
    SnapshotData small(HOST_PAGE_SIZE, 4 * HOST_PAGE_SIZE);
    auto mem = allocatePrivateMemory(2 * HOST_PAGE_SIZE);
    mem[HOST_PAGE_SIZE + 7] = 5;
    small.fillGapsWithBytewiseRegions();
    auto diffs = small.diffWithDirtyRegions(std::span<uint8_t>(mem.get(), 2 * HOST_PAGE_SIZE), { 0, 1 });
    REQUIRE_EQ(diffs.size(), 1u);
    REQUIRE_EQ(diffs[0].getOffset(), (uint64_t)HOST_PAGE_SIZE);
    REQUIRE_EQ(diffs[0].getData().size(), (size_t)HOST_PAGE_SIZE);
}

TEST_CASE("diffArrayRegions chunks", "[util][snapshot]")
{
    std::vector<uint8_t> a(1024, 0), b(1024, 0);
    b[5] = 1;
    b[6] = 1;
    b[300] = 1;
    b[1023] = 1;
    std::vector<std::pair<uint64_t, uint64_t>> regs;
    diffArrayRegions(regs, 0, 1024, a, b);
    REQUIRE_EQ(regs.size(), 3u);
    REQUIRE_EQ(regs[0].first, 5u);
    REQUIRE_EQ(regs[0].second, 2u);
    REQUIRE_EQ(regs[1].first, 300u);
    REQUIRE_EQ(regs[2].first, 1023u);
    REQUIRE_EQ(regs[2].second, 1u);
}

TEST_CASE("delta encode / apply", "[util]")
{
    std::vector<uint8_t> oldData(20000, 0), newData(24000, 0);
    for (size_t i = 0; i < oldData.size(); i++) {
        oldData[i] = (uint8_t)(i * 7);
        newData[i] = oldData[i];
    }
    newData[5] = 99;
    newData[12000] = 98;
    newData[23999] = 97;
    for (std::string def : { "pages=4096;xor;zstd=1", "pages=64;", "xor;", "" }) {
        DeltaSettings cfg(def);
        if (cfg.useZstd && !deltaZstdAvailable()) {
            cfg.useZstd = false;
        }
        auto delta = serializeDelta(cfg, oldData.data(), oldData.size(), newData.data(), newData.size());
        std::vector<uint8_t> work = oldData;
        applyDelta(
          delta, [&](uint32_t sz) { work.resize(sz); }, [&] { return work.data(); });
        REQUIRE(work == newData);
        if (cfg.usePages) {
            REQUIRE(delta.size() < newData.size());
        }
    }
}

TEST_CASE("random strings and clock", "[util]")
{
    auto a = randomString(16);
    auto b = randomString(16);
    REQUIRE_EQ(a.size(), 16u);
    REQUIRE(a != b);
    auto& clock = getGlobalClock();
    auto t0 = clock.now();
    SLEEP_MS(15);
    REQUIRE(clock.timeDiff(clock.now(), t0) >= 10);
    REQUIRE(clock.epochMillis() > 1600000000000L);
}
