// Utilities without a dedicated case elsewhere (strategy: reference
// tests/test/util/test_{periodic_thread,tokens,environment,network,memory,
// batch,func}.cpp)
#include "harness.h"

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/PeriodicBackgroundThread.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/environment.h>
#include <faabric/util/func.h>
#include <faabric/util/locks.h>
#include <faabric/util/memory.h>
#include <faabric/util/network.h>
#include <faabric/util/queue.h>

#include <atomic>
#include <cstring>
#include <thread>
#include <unistd.h>

using namespace faabric::util;

namespace {
class CountingThread : public PeriodicBackgroundThread
{
  public:
    std::atomic<int> ticks{ 0 };
    std::atomic<int> tidied{ 0 };

    void doWork() override { ticks++; }

    void tidyUp() override { tidied++; }
};
}

TEST_CASE("periodic background thread ticks, stops promptly and tidies up", "[util]")
{
    CountingThread t;
    t.startMs(5);
    for (int i = 0; i < 400 && t.ticks.load() < 3; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    REQUIRE(t.ticks.load() >= 3);
    auto t0 = std::chrono::steady_clock::now();
    t.stop();
    auto stopMs = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    REQUIRE(stopMs < 1000);
    REQUIRE_EQ(t.tidied.load(), 1);
    int after = t.ticks.load();
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    REQUIRE_EQ(t.ticks.load(), after);
    // stopping twice / never-started threads is harmless
    t.stop();
    CountingThread idle;
    idle.stop();
    // a long interval does not delay shutdown
    CountingThread slow;
    slow.start(3600);
    REQUIRE_EQ(slow.getIntervalSeconds(), 3600);
    t0 = std::chrono::steady_clock::now();
    slow.stop();
    stopMs = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    REQUIRE(stopMs < 1000);
    REQUIRE_EQ(slow.ticks.load(), 0);
}

TEST_CASE("token pool hands out each token once and blocks when empty", "[util]")
{
    TokenPool pool(3);
    REQUIRE_EQ(pool.size(), 3);
    REQUIRE_EQ(pool.free(), 3);
    std::set<int> got = { pool.getToken(), pool.getToken(), pool.getToken() };
    REQUIRE_EQ(got.size(), (size_t)3);
    REQUIRE_EQ(pool.taken(), 3);
    REQUIRE_EQ(pool.free(), 0);
    // a fourth taker waits for a release
    std::atomic<int> late{ -1 };
    std::thread waiter([&] { late = pool.getToken(); });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    REQUIRE_EQ(late.load(), -1);
    pool.releaseToken(*got.begin());
    waiter.join();
    REQUIRE_EQ(late.load(), *got.begin());
    pool.reset();
    REQUIRE_EQ(pool.free(), 3);
    // a pool of zero never blocks: it hands out a dummy token
    TokenPool none(0);
    REQUIRE_EQ(none.getToken(), -1);
}

TEST_CASE("environment variables and core counts", "[util]")
{
    REQUIRE_EQ(getEnvVar("FB_TEST_SURELY_UNSET", "fallback"), std::string("fallback"));
    setEnvVar("FB_TEST_VAR", "abc");
    REQUIRE_EQ(getEnvVar("FB_TEST_VAR", "fallback"), std::string("abc"));
    // empty counts as unset
    setEnvVar("FB_TEST_VAR", "");
    REQUIRE_EQ(getEnvVar("FB_TEST_VAR", "fallback"), std::string("fallback"));
    unsetEnvVar("FB_TEST_VAR");
    REQUIRE_EQ(getEnvVar("FB_TEST_VAR", "gone"), std::string("gone"));

    unsigned int real = getUsableCores();
    REQUIRE(real >= 1);
    auto& conf = getSystemConfig();
    int before = conf.overrideCpuCount;
    conf.overrideCpuCount = 3;
    REQUIRE_EQ(getUsableCores(), 3u);
    conf.overrideCpuCount = before;
    REQUIRE_EQ(getUsableCores(), real);
    REQUIRE(getUsableGpus() >= 0);
}

TEST_CASE("network helpers resolve names and this host's address", "[util]")
{
    REQUIRE_EQ(getIPFromHostname("localhost"), std::string("127.0.0.1"));
    REQUIRE_EQ(getIPFromHostname("127.0.0.1"), std::string("127.0.0.1"));
    // unknown names come back empty (callers fall back to the name itself)
    REQUIRE_EQ(getIPFromHostname("no-such-host.invalid"), std::string(""));
    std::string ip = getPrimaryIPForThisHost("");
    REQUIRE(!ip.empty());
    REQUIRE_EQ(std::count(ip.begin(), ip.end(), '.'), 3L);
    REQUIRE_EQ(getPrimaryIPForThisHost("lo"), std::string("127.0.0.1"));
}

TEST_CASE("memory regions: fds, shared and private mappings", "[util][memory]")
{
    const size_t page = HOST_PAGE_SIZE;
    REQUIRE(isPageAligned((void*)(uintptr_t)(4 * page)));
    REQUIRE(!isPageAligned((void*)(uintptr_t)(4 * page + 8)));

    // fresh memory is zeroed, writable and page aligned
    MemoryRegion priv = allocatePrivateMemory(3 * page);
    MemoryRegion shared = allocateSharedMemory(3 * page);
    REQUIRE(isPageAligned(priv.get()) && isPageAligned(shared.get()));
    REQUIRE(priv[0] == 0 && priv[3 * page - 1] == 0);
    priv[10] = 1;
    shared[10] = 2;

    // a memfd backs two views: shared ones see each other's writes, a private
    // one keeps its own copy after the first write
    int fd = createFd(2 * page, "fb-test");
    REQUIRE(fd > 0);
    std::vector<uint8_t> pattern(2 * page);
    for (size_t i = 0; i < pattern.size(); i++) {
        pattern[i] = (uint8_t)(i % 251);
    }
    writeToFd(fd, 0, pattern);
    MemoryRegion viewA = allocatePrivateMemory(2 * page);
    MemoryRegion viewB = allocatePrivateMemory(2 * page);
    MemoryRegion viewC = allocatePrivateMemory(2 * page);
    mapMemoryShared({ viewA.get(), 2 * page }, fd);
    mapMemoryShared({ viewB.get(), 2 * page }, fd);
    mapMemoryPrivate({ viewC.get(), 2 * page }, fd);
    REQUIRE(memcmp(viewA.get(), pattern.data(), 2 * page) == 0);
    REQUIRE(memcmp(viewC.get(), pattern.data(), 2 * page) == 0);
    viewA[5] = 200;
    REQUIRE_EQ(viewB[5], 200);
    REQUIRE_EQ(viewC[5], 200); // untouched private pages still follow the file
    viewC[6] = 77;             // copy-on-write from here on
    REQUIRE_EQ(viewA[6], pattern[6]);
    viewA[7] = 201;
    REQUIRE_EQ(viewC[7], pattern[7]);

    // growing the file and appending
    resizeFd(fd, 3 * page);
    std::vector<uint8_t> extra(page, 9);
    writeToFd(fd, (off_t)(2 * page), extra);
    MemoryRegion grown = allocatePrivateMemory(3 * page);
    mapMemoryShared({ grown.get(), 3 * page }, fd);
    REQUIRE_EQ(grown[2 * page + 17], 9);
    std::vector<uint8_t> tail(page, 4);
    appendDataToFd(fd, tail);
    MemoryRegion appended = allocatePrivateMemory(4 * page);
    mapMemoryShared({ appended.get(), 4 * page }, fd);
    REQUIRE_EQ(appended[3 * page + 1], 4);
    ::close(fd);

    // misaligned targets and bad fds are refused
    REQUIRE_THROWS(mapMemoryShared({ viewA.get() + 8, page }, fd));
    int other = createFd(page, "fb-test-2");
    REQUIRE_THROWS(mapMemoryShared({ viewA.get() + 8, page }, other));
    ::close(other);
    REQUIRE_THROWS(mapMemoryPrivate({ viewA.get(), page }, -1));

    // reserve-then-claim
    MemoryRegion reserved = allocateVirtualMemory(8 * page);
    claimVirtualMemory({ reserved.get(), 2 * page });
    reserved[2 * page - 1] = 5;
    REQUIRE_EQ(reserved[2 * page - 1], 5);
}

TEST_CASE("batch helpers: validity, ids and status objects", "[util]")
{
    auto ber = batchExecFactory("demo", "echo", 3);
    REQUIRE(isBatchExecRequestValid(ber));
    REQUIRE_EQ(ber->messages_size(), 3);
    for (const auto& m : ber->messages()) {
        REQUIRE_EQ(m.appid(), ber->appid());
        REQUIRE_EQ(m.user(), std::string("demo"));
    }
    REQUIRE(!isBatchExecRequestValid(nullptr));
    // a message for another user invalidates the request; another (non-empty)
    // function name does not: requests may carry calls chained by name
    // (reference: src/util/batch.cpp:58-78)
    auto chainedByName = batchExecFactory("demo", "echo", 2);
    chainedByName->mutable_messages(1)->set_function("other");
    REQUIRE(isBatchExecRequestValid(chainedByName));
    auto broken = batchExecFactory("demo", "echo", 2);
    broken->mutable_messages(1)->set_user("somebody-else");
    REQUIRE(!isBatchExecRequestValid(broken));
    auto noFunction = batchExecFactory("demo", "echo", 2);
    noFunction->mutable_messages(0)->set_function("");
    REQUIRE(!isBatchExecRequestValid(noFunction));
    auto wrongApp = batchExecFactory("demo", "echo", 2);
    wrongApp->mutable_messages(0)->set_appid(wrongApp->appid() + 1);
    REQUIRE(!isBatchExecRequestValid(wrongApp));
    REQUIRE(!isBatchExecRequestValid(batchExecFactory()));

    updateBatchExecAppId(ber, 4321);
    updateBatchExecGroupId(ber, 8765);
    REQUIRE_EQ(ber->appid(), 4321);
    REQUIRE_EQ(ber->groupid(), 8765);
    for (const auto& m : ber->messages()) {
        REQUIRE_EQ(m.appid(), 4321);
        REQUIRE_EQ(m.groupid(), 8765);
    }
    REQUIRE(isBatchExecRequestValid(ber));

    auto status = batchExecStatusFactory(ber);
    REQUIRE_EQ(status->appid(), 4321);
    REQUIRE_EQ(status->expectednummessages(), 3);
    REQUIRE(!status->finished());
    REQUIRE_EQ(batchExecStatusFactory(99)->appid(), 99);
}

TEST_CASE("function helpers: names, keys and async responses", "[util]")
{
    auto msg = messageFactory("demo", "echo");
    REQUIRE(msg.id() > 0);
    REQUIRE_EQ(funcToString(msg, false), std::string("demo/echo"));
    REQUIRE_EQ(funcToString(msg, true), "demo/echo:" + std::to_string(msg.id()));
    auto ber = batchExecFactory("demo", "echo", 2);
    REQUIRE_EQ(funcToString(ber), "demo/echo:" + std::to_string(ber->appid()));

    // ids: kept when present, generated (with the derived keys) when missing
    unsigned int original = msg.id();
    REQUIRE_EQ(setMessageId(msg), original);
    faabric::Message blank;
    unsigned int fresh = setMessageId(blank);
    REQUIRE(fresh > 0);
    REQUIRE_EQ(blank.id(), (int)fresh);
    REQUIRE_EQ(blank.resultkey(), resultKeyFromMessageId(fresh));
    REQUIRE_EQ(blank.statuskey(), statusKeyFromMessageId(fresh));
    REQUIRE(resultKeyFromMessageId(7) != statusKeyFromMessageId(7));
    REQUIRE_EQ(buildAsyncResponse(msg), std::to_string(msg.id()));

    // every thread of an app shares one main-thread snapshot key
    auto other = messageFactory("demo", "echo");
    other.set_appid(msg.appid());
    REQUIRE_EQ(getMainThreadSnapshotKey(msg), getMainThreadSnapshotKey(other));
    other.set_appid(msg.appid() + 1);
    REQUIRE(getMainThreadSnapshotKey(msg) != getMainThreadSnapshotKey(other));

    msg.set_cmdline("prog --alpha 1  beta");
    auto argv = getArgvForMessage(msg);
    REQUIRE(argv == (std::vector<std::string>{ "function.wasm", "prog", "--alpha", "1", "beta" }));

    auto shared = messageFactoryShared("demo", "x");
    REQUIRE(shared != nullptr && shared->function() == "x");
    // wire round trip
    faabric::Message parsed;
    auto bytes = messageToBytes(msg);
    REQUIRE(parsed.ParseFromArray(bytes.data(), (int)bytes.size()));
    REQUIRE_EQ(parsed.id(), msg.id());
    REQUIRE_EQ(parsed.cmdline(), msg.cmdline());
}

// ---------------------------------------------------------------------------
// Timers, logging, config dump, exec-graph helpers, crash handler
// ---------------------------------------------------------------------------
#include <faabric/util/ExecGraph.h>
#include <faabric/util/crash.h>
#include <faabric/util/files.h>
#include <faabric/util/logging.h>
#include <faabric/util/timing.h>

#include <filesystem>

TEST_CASE("timers accumulate per label and print a sorted table", "[util]")
{
    clearTimerTotals();
    startGlobalTimer();
    auto t = startTimer();
    std::this_thread::sleep_for(std::chrono::milliseconds(3));
    REQUIRE(getTimeDiffNanos(t) >= 3'000'000L);
    REQUIRE(getTimeDiffMicros(t) >= 3000L);
    REQUIRE(getTimeDiffMillis(t) >= 3.0);
    logEndTimer("slow", t);
    auto quick = startTimer();
    logEndTimer("quick", quick);
    logEndTimer("quick", quick);
    std::string totals = getTimerTotalsString();
    // "label:totalMicros:count" lines, largest total first
    REQUIRE(totals.find("slow:") < totals.find("quick:"));
    REQUIRE(totals.find("quick:") != std::string::npos);
    REQUIRE(totals.substr(totals.find("quick:")).find(":2") != std::string::npos);
    printTimerTotals();
    clearTimerTotals();
    REQUIRE(getTimerTotalsString().find("slow") == std::string::npos);

    timespec ts{ 3, 500 };
    REQUIRE_EQ(timespecToNanos(&ts), (uint64_t)3'000'000'500ULL);
    timespec back{};
    nanosToTimespec(3'000'000'500ULL, &back);
    REQUIRE(back.tv_sec == 3 && back.tv_nsec == 500);
}

TEST_CASE("logging levels filter, names parse, LOG_FILE redirects", "[util]")
{
    LogLevel before = getLogLevel();
    setLogLevel("debug");
    REQUIRE(getLogLevel() == LogLevel::debug);
    setLogLevel("warn");
    REQUIRE(getLogLevel() == LogLevel::warn);
    setLogLevel("off");
    REQUIRE(getLogLevel() == LogLevel::off);
    setLogLevel(LogLevel::err);
    REQUIRE(getLogLevel() == LogLevel::err);
    // unknown names mean the default
    setLogLevel("shouting");
    REQUIRE(getLogLevel() == LogLevel::info);

    // LOG_FILE: lines land in the file with the reference's pattern
    const std::string path = "/tmp/fb_log_" + std::to_string(getpid()) + ".log";
    ::unlink(path.c_str());
    auto& conf = getSystemConfig();
    std::string oldFile = conf.logFile, oldLevel = conf.logLevel;
    conf.logFile = path;
    conf.logLevel = "info";
    initLogging();
    SPDLOG_INFO("hello {} and {}", 42, "world");
    SPDLOG_DEBUG("filtered {}", 1);
    conf.logFile = oldFile;
    conf.logLevel = oldLevel;
    initLogging();
    std::string contents = readFileToString(path);
    REQUIRE(contents.find("hello 42 and world") != std::string::npos);
    REQUIRE(contents.find("[I]") != std::string::npos);
    REQUIRE(contents.find("filtered") == std::string::npos);
    ::unlink(path.c_str());
    setLogLevel(before);
}

TEST_CASE("config dump lists every knob", "[util]")
{
    auto& conf = getSystemConfig();
    LogLevel before = getLogLevel();
    const std::string path = "/tmp/fb_conf_" + std::to_string(getpid()) + ".log";
    std::string oldFile = conf.logFile;
    conf.logFile = path;
    initLogging();
    setLogLevel(LogLevel::info);
    conf.print();
    conf.logFile = oldFile;
    initLogging();
    setLogLevel(before);
    std::string dump = readFileToString(path);
    ::unlink(path.c_str());
    for (const char* knob : { "LOG_LEVEL", "BATCH_SCHEDULER_MODE", "GLOBAL_MESSAGE_TIMEOUT", "DIRTY_TRACKING_MODE", "PLANNER_HOST", "FAABRIC_GPUS", "FAABRIC_CHECKPOINT_DIR" }) {
        REQUIRE(dump.find(knob) != std::string::npos);
    }
}

TEST_CASE("exec graph helpers: counting, hosts, MPI ranks, JSON, details", "[util]")
{
    auto mk = [](int id, const std::string& host, int rank = -1) {
        faabric::Message m = messageFactory("demo", "node");
        m.set_id(id);
        m.set_executedhost(host);
        if (rank >= 0) {
            m.set_ismpi(true);
            m.set_mpirank(rank);
            m.set_mpiworldsize(3);
        }
        return m;
    };
    ExecGraphNode leafA{ mk(2, "hostB", 1), {} };
    ExecGraphNode leafB{ mk(3, "hostA", 2), {} };
    ExecGraphNode root{ mk(1, "hostA", 0), { leafA, leafB } };
    ExecGraph graph{ root };
    REQUIRE_EQ(countExecGraphNodes(graph), 3);
    REQUIRE(getExecGraphHosts(graph) == (std::set<std::string>{ "hostA", "hostB" }));
    REQUIRE(getMpiRankHostsFromExecGraph(graph) == (std::vector<std::string>{ "hostA", "hostB", "hostA" }));
    std::string json = execGraphToJson(graph);
    REQUIRE(json.find("\"root\"") != std::string::npos);
    REQUIRE(json.find("\"chained\"") != std::string::npos);
    REQUIRE(json.find("hostB") != std::string::npos);
    REQUIRE(execNodeToJson(leafA).find("\"msg\"") != std::string::npos);

    // chained-call bookkeeping and detail counters only when recording
    faabric::Message parent = messageFactory("demo", "parent");
    faabric::Message child = messageFactory("demo", "child");
    logChainedFunction(parent, child);
    REQUIRE(getChainedFunctions(parent) == (std::set<unsigned int>{ (unsigned int)child.id() }));
    addDetail(parent, "note", "ignored");
    incrementCounter(parent, "hits");
    REQUIRE_EQ(parent.execgraphdetails_size(), 0);
    REQUIRE_EQ(parent.intexecgraphdetails_size(), 0);
    parent.set_recordexecgraph(true);
    addDetail(parent, "note", "kept");
    incrementCounter(parent, "hits");
    incrementCounter(parent, "hits", 4);
    REQUIRE_EQ(parent.execgraphdetails().at("note"), std::string("kept"));
    REQUIRE_EQ(parent.intexecgraphdetails().at("hits"), 5);
}

TEST_CASE("crash handler prints a stack trace for the test signal", "[util]")
{
    // The real signals re-raise; the test signal only reports
    setUpCrashHandler(12341234);
    printStackTrace(nullptr);
    REQUIRE(true);
}

// ---------------------------------------------------------------------------
// Wire codec and JSON layer under hostile input
// ---------------------------------------------------------------------------
#include <faabric/proto/wire.h>
#include <faabric/util/json.h>

TEST_CASE("wire reader: skipping, truncation and overlong varints", "[proto]")
{
    using faabric::proto::Reader;
    using faabric::proto::Writer;
    Writer w;
    w.varint(1, 300, false);
    w.varint(2, 0, false); // default: omitted
    w.varint(3, 0, true);  // forced
    w.str(4, "payload", false);
    w.str(5, "", false);
    std::string buf = w.take();
    Reader r(buf);
    uint32_t field = 0;
    int wt = -1;
    std::vector<uint32_t> seen;
    while (r.next(field, wt)) {
        seen.push_back(field);
        REQUIRE(r.skip(wt));
    }
    REQUIRE(r.ok());
    REQUIRE(seen == (std::vector<uint32_t>{ 1, 3, 4 }));

    // fixed-width fields of unknown messages are skipped too
    std::string fixed = { 0x0d, 1, 2, 3, 4, 0x11, 1, 2, 3, 4, 5, 6, 7, 8, 0x18, 0x05 };
    Reader rf(fixed);
    int n = 0;
    while (rf.next(field, wt)) {
        REQUIRE(rf.skip(wt));
        n++;
    }
    REQUIRE(rf.ok());
    REQUIRE_EQ(n, 3);

    // every prefix of a valid message either parses or fails cleanly
    faabric::Message m = messageFactory("demo", "echo");
    m.set_inputdata(std::string(300, 'x'));
    m.set_cmdline("a b c");
    (*m.mutable_intexecgraphdetails())["k"] = 7;
    m.add_chainedmsgids(9);
    std::string wire = m.SerializeAsString();
    int parsedOk = 0;
    for (size_t cut = 0; cut <= wire.size(); cut++) {
        faabric::Message p;
        if (p.ParseFromArray(wire.data(), (int)cut)) {
            parsedOk++;
        }
    }
    REQUIRE(parsedOk >= 2); // the empty prefix and the whole message at least
    faabric::Message whole;
    REQUIRE(whole.ParseFromString(wire));
    REQUIRE_EQ(whole.intexecgraphdetails().at("k"), 7);
    REQUIRE_EQ(whole.chainedmsgids(0), 9u);

    // an 11-byte varint and a group wire type are rejected
    std::string overlong(11, (char)0x80);
    overlong.insert(overlong.begin(), 0x08);
    faabric::Message p;
    REQUIRE(!p.ParseFromString(overlong));
    REQUIRE(!p.ParseFromString(std::string({ 0x0b })));
    // random bytes never crash the parser
    uint32_t seed = 12345;
    for (int i = 0; i < 2000; i++) {
        std::string junk;
        int len = (int)(seed % 64);
        for (int j = 0; j < len; j++) {
            seed = seed * 1664525u + 1013904223u;
            junk.push_back((char)(seed >> 24));
        }
        faabric::BatchExecuteRequest ber;
        (void)ber.ParseFromString(junk);
        seed = seed * 1664525u + 1013904223u;
    }
}

TEST_CASE("json layer: values, escapes, errors and schema names", "[proto]")
{
    using faabric::proto::JsonValue;
    JsonValue v = JsonValue::parse(R"({"a": [1, -2.5, "3", true, null], "b": {"nested": "x\n\"y\" é"}, "big": "9007199254740993"})");
    REQUIRE(v.isObject());
    const JsonValue* a = v.find("a");
    REQUIRE(a != nullptr && a->isArray() && a->elements().size() == 5);
    REQUIRE_EQ(a->elements()[0].asInt(), (int64_t)1);
    REQUIRE_EQ(a->elements()[1].asDouble(), -2.5);
    REQUIRE_EQ(a->elements()[2].asInt(), (int64_t)3); // numbers as strings
    REQUIRE(a->elements()[3].asBool());
    REQUIRE(a->elements()[4].isNull());
    REQUIRE(v.find("b")->find("nested")->asString().find("\"y\"") != std::string::npos);
    REQUIRE_EQ(v.find("big")->asInt(), (int64_t)9007199254740993LL);
    REQUIRE(v.find("missing") == nullptr);
    for (const char* bad : { "{", "{\"a\": }", "[1, 2", "{\"a\" 1}", "tru", "\"unterminated", "{\"a\": 1} trailing", "" }) {
        REQUIRE_THROWS(JsonValue::parse(bad));
    }

    // base64 for bytes fields, both directions, all paddings
    for (const std::string& raw : { std::string(""), std::string("f"), std::string("fo"), std::string("foo"), std::string("\x00\xff\x10", 3) }) {
        REQUIRE_EQ(faabric::proto::base64Decode(faabric::proto::base64Encode(raw)), raw);
    }
    REQUIRE_EQ(faabric::proto::base64Encode("foob"), std::string("Zm9vYg=="));

    // messages use the schema's json names and survive a round trip
    faabric::Message m = messageFactory("demo", "echo");
    m.set_inputdata(std::string("\x01\x02\x03", 3));
    m.set_outputdata("line\nbreak \"quoted\"");
    m.set_mpiworldsize(4);
    m.set_ismpi(true);
    std::string json = messageToJson(m);
    REQUIRE(json.find("\"input_data\"") != std::string::npos);
    REQUIRE(json.find("\"mpi_world_size\"") != std::string::npos);
    faabric::Message back;
    jsonToMessage(json, &back);
    REQUIRE_EQ(back.inputdata(), m.inputdata());
    REQUIRE_EQ(back.outputdata(), m.outputdata());
    REQUIRE_EQ(back.mpiworldsize(), 4);
    REQUIRE(back.ismpi());
    // unknown keys are ignored, wrong shapes are errors
    faabric::Message lenient;
    jsonToMessage(R"({"id": 5, "no_such_field": [1, 2, 3]})", &lenient);
    REQUIRE_EQ(lenient.id(), 5);
    REQUIRE_THROWS(jsonToMessage("[1, 2]", &lenient));
    REQUIRE_THROWS(jsonToMessage("not json", &lenient));
}

// ---------------------------------------------------------------------------
// CPU pinning / GPU placement helpers
// ---------------------------------------------------------------------------
#include <faabric/util/hwloc.h>

#include <sched.h>

TEST_CASE("cpu pinning: claims are exclusive, released and exhaustible", "[util]")
{
    auto& conf = getSystemConfig();
    const int before = conf.overrideCpuCount;
    const int free0 = getNumFreeCpus();
    REQUIRE(free0 >= 1);
    std::set<int> cpus;
    {
        std::vector<std::unique_ptr<FaabricCpuSet>> pins;
        // each thread gets a CPU of its own and really runs there
        std::mutex mx;
        std::vector<std::thread> ts;
        const int n = std::min(free0, 3);
        std::atomic<int> wrongCpu{ 0 };
        for (int i = 0; i < n; i++) {
            ts.emplace_back([&] {
                auto pin = pinThreadNearGpu(pthread_self(), 0); // no GPU here: any free CPU
                // (assertions stay on the main thread: the harness counts them
                // without synchronisation)
                if (sched_getcpu() != pin->getCpuIdx() || !CPU_ISSET(pin->getCpuIdx(), pin->get()) ||
                    CPU_COUNT(pin->get()) != 1) {
                    wrongCpu++;
                }
                std::lock_guard<std::mutex> lk(mx);
                cpus.insert(pin->getCpuIdx());
                pins.push_back(std::move(pin));
            });
        }
        for (auto& t : ts) {
            t.join();
        }
        REQUIRE_EQ(wrongCpu.load(), 0);
        REQUIRE_EQ((int)cpus.size(), n);
        REQUIRE_EQ(getNumFreeCpus(), free0 - n);
        // claim the rest, then one more
        std::atomic<bool> exhaustedThrew{ false };
        std::thread rest([&] {
            while (getNumFreeCpus() > 0) {
                pins.push_back(pinThreadToFreeCpu(pthread_self()));
            }
            // (outside test mode: in test mode an extra pin shares CPU 0)
            setTestMode(false);
            try {
                pinThreadToFreeCpu(pthread_self());
            } catch (const std::runtime_error&) {
                exhaustedThrew = true;
            }
            setTestMode(true);
        });
        rest.join();
        REQUIRE(exhaustedThrew.load());
        REQUIRE_EQ(getNumFreeCpus(), 0);
    }
    // RAII: everything is free again
    REQUIRE_EQ(getNumFreeCpus(), free0);
    conf.overrideCpuCount = before;

    // placement: no GPU -> -1, and binding is a no-op rather than an error
    if (getUsableGpus() == 0) {
        REQUIRE_EQ(gpuForRank(0), -1);
        REQUIRE_EQ(gpuForRank(5), -1);
    } else {
        REQUIRE_EQ(gpuForRank(0), 0);
        REQUIRE_EQ(gpuForRank(getUsableGpus()), 0);
    }
    bindThreadToGpu(-1);
    bindThreadToGpu(0);
}

TEST_CASE("delta: the building blocks produce and walk the same stream as the one-shot codec", "[util][delta]")
{
    std::vector<uint8_t> oldData(3 * 4096, 7), newData(3 * 4096, 7);
    for (int i = 4096; i < 4096 + 64; i++) {
        newData[i] = (uint8_t)i;
    }
    for (const char* def : { "pages=4096;xor;", "pages=4096;", "pages=4096;xor;zstd=1;" }) {
        faabric::util::DeltaSettings cfg(def);
        if (cfg.useZstd && !faabric::util::deltaZstdAvailable()) {
            continue;
        }
        auto whole = faabric::util::serializeDelta(cfg, oldData.data(), oldData.size(), newData.data(), newData.size());
        // by hand: one run covering the changed page
        std::vector<uint8_t> payload(newData.begin() + 4096, newData.begin() + 8192);
        if (cfg.xorWithOld) {
            for (size_t i = 0; i < payload.size(); i++) {
                payload[i] ^= oldData[4096 + i];
            }
        }
        std::vector<uint8_t> cmds;
        faabric::util::deltaBegin(cmds, (uint32_t)newData.size());
        faabric::util::deltaAppendRun(cmds, cfg.xorWithOld, 4096, payload.data(), (uint32_t)payload.size());
        faabric::util::deltaAppendRun(cmds, cfg.xorWithOld, 0, payload.data(), 0); // empty runs are dropped
        REQUIRE(faabric::util::deltaFinish(cfg, std::move(cmds)) == whole);
        // walking it reports the size and exactly that run
        uint32_t total = 0;
        int runs = 0;
        faabric::util::deltaForEach(
          whole,
          [&](uint32_t t) { total = t; },
          [&](bool isXor, uint32_t offset, const uint8_t* p, uint32_t length) {
              runs++;
              REQUIRE_EQ(isXor, cfg.xorWithOld);
              REQUIRE_EQ(offset, 4096u);
              REQUIRE_EQ(length, 4096u);
              REQUIRE(memcmp(p, payload.data(), length) == 0);
          });
        REQUIRE_EQ(total, (uint32_t)newData.size());
        REQUIRE_EQ(runs, 1);
    }
    // truncated and unknown commands are refused
    std::vector<uint8_t> bad = { faabric::util::DELTACMD_DELTA_XOR, 0, 0, 0, 0, 9, 0, 0, 0, 1 };
    REQUIRE_THROWS(faabric::util::deltaForEach(bad, [](uint32_t) {}, [](bool, uint32_t, const uint8_t*, uint32_t) {}));
    std::vector<uint8_t> unknown = { 0x77 };
    REQUIRE_THROWS(faabric::util::deltaForEach(unknown, [](uint32_t) {}, [](bool, uint32_t, const uint8_t*, uint32_t) {}));
}

TEST_CASE("bytes, strings and batch helpers under the reference's names", "[util]")
{
    std::vector<uint8_t> four;
    faabric::util::appendBytesOf<int>(four, 0x01020304);
    REQUIRE_EQ(four.size(), sizeof(int));
    REQUIRE_EQ(faabric::util::bytesToInt(four), 0x01020304);
    faabric::util::appendBytesOf<uint8_t>(four, 9);
    REQUIRE_EQ(four.size(), 5u);
    REQUIRE_THROWS(faabric::util::bytesToInt(four));
    REQUIRE_EQ(faabric::util::intToHexString<uint8_t>(0x0f), std::string("0f"));
    REQUIRE_EQ(faabric::util::intToHexString<uint16_t>(0xabc), std::string("0abc"));
    REQUIRE_EQ(faabric::util::intToHexString<uint32_t>(255), std::string("000000ff"));
    REQUIRE_EQ(faabric::util::intToHexString<int>(-1), std::string("ffffffff"));
    REQUIRE_EQ(faabric::util::vectorToString<int>({ 1, 2, 3 }), std::string("[1, 2, 3]"));
    REQUIRE_EQ(faabric::util::vectorToString<std::string>({ "a", "bc" }), std::string("[a, bc]"));
    REQUIRE_EQ(faabric::util::vectorToString<int>({}), std::string("[]"));
    auto status = faabric::util::batchExecStatusFactory(7);
    for (int rv : { 0, 1, MIGRATED_FUNCTION_RETURN_VALUE, 0 }) {
        status->add_messageresults()->set_returnvalue(rv);
    }
    REQUIRE_EQ(faabric::util::getNumFinishedMessagesInBatch(status), 3);
}
