// Strings, exec graphs, the concurrent map, periodic threads, protobuf-style
// messages and the dirty-tracker matrix, one case per case / section of the
// reference's suites
// (reference: tests/test/util/test_strings.cpp, test_exec_graph.cpp,
// test_concurrent_map.cpp, test_periodic_thread.cpp, test_state.cpp,
// test_dirty.cpp, tests/test/proto/test_proto.cpp)
#include "fixtures.h"

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/PeriodicBackgroundThread.h>
#include <faabric/util/barrier.h>
#include <faabric/util/bytes.h>
#include <faabric/util/concurrent_map.h>
#include <faabric/util/dirty.h>
#include <faabric/util/memory.h>
#include <faabric/util/state.h>
#include <faabric/util/string_tools.h>

#include <thread>

using namespace faabric::util;
using namespace tests;

// ---- strings -----------------------------------------------------------------
TEST_CASE("string case: bytes to string and back", "[util][strings][cases]")
{
    std::string in = "abcdefghijkl12345";
    auto bytes = stringToBytes(in);
    REQUIRE_EQ(bytes.size(), in.size());
    std::string out = bytesToString(bytes);
    REQUIRE_EQ(out.size(), in.size());
    REQUIRE_EQ(out, in);
}

TEST_CASE("string case: all-whitespace", "[util][strings][cases]")
{
    REQUIRE(isAllWhitespace("    "));
    REQUIRE(!isAllWhitespace("  s  "));
}

TEST_CASE("string case: startsWith (an empty prefix does not count)", "[util][strings][cases]")
{
    REQUIRE(startsWith("foobar", "foo"));
    REQUIRE(!startsWith("foobar", "goo"));
    REQUIRE(!startsWith("foobar", ""));
}

TEST_CASE("string case: endsWith (an empty suffix does not count)", "[util][strings][cases]")
{
    REQUIRE(endsWith("foobar", "bar"));
    REQUIRE(!endsWith("foobar", "foo"));
    REQUIRE(!endsWith("foobar", "ob"));
    REQUIRE(!endsWith("foobar", ""));
    REQUIRE(!endsWith("", "foobar"));
}

TEST_CASE("string case: removing a substring", "[util][strings][cases]")
{
    REQUIRE_EQ(removeSubstr("blah foobar", "blah"), std::string(" foobar"));
    REQUIRE_EQ(removeSubstr("blahblah", ""), std::string("blahblah"));
    REQUIRE_EQ(removeSubstr("", "foobar"), std::string(""));
    REQUIRE_EQ(removeSubstr("foo bar baz", "bar"), std::string("foo  baz"));
}

TEST_CASE("string case: which strings are integers", "[util][strings][cases]")
{
    REQUIRE(stringIsInt("12345"));
    REQUIRE(stringIsInt("0"));
    REQUIRE(!stringIsInt(" 12345"));
    REQUIRE(!stringIsInt("123 "));
    REQUIRE(!stringIsInt("abcd"));
    REQUIRE(!stringIsInt("12a33"));
}

TEST_CASE("string case: a vector of ints as text", "[util][strings][cases]")
{
    REQUIRE_EQ(vectorToString<int>({ -1, 1, -2, 3 }), std::string("[-1, 1, -2, 3]"));
}

TEST_CASE("string case: a vector of strings as text", "[util][strings][cases]")
{
    REQUIRE_EQ(vectorToString<std::string>({ "foo", "blah", "baz" }), std::string("[foo, blah, baz]"));
}

TEST_CASE("state key case: user and key joined by an underscore", "[util][state][cases]")
{
    REQUIRE_EQ(keyForUser("foo", "bar"), std::string("foo_bar"));
}

// ---- exec graph ----------------------------------------------------------------
namespace {
bool sameGraph(const ExecGraphNode& a, const ExecGraphNode& b)
{
    if (a.msg.id() != b.msg.id() || a.children.size() != b.children.size()) {
        return false;
    }
    for (size_t i = 0; i < a.children.size(); i++) {
        if (!sameGraph(a.children[i], b.children[i])) {
            return false;
        }
    }
    return true;
}
}

TEST_CASE("exec graph case: a three-level chain of seven calls is rebuilt from their results", "[util][exec-graph][cases]")
{
    const int nMsg = 7;
    ClusterFixture f(nMsg);
    auto ber = batchExecFactory("demo", "echo", nMsg);
    std::vector<faabric::Message> m;
    for (int i = 0; i < nMsg; i++) {
        m.push_back(ber->messages(i));
        m.back().set_executedhost(f.conf.endpointHost);
    }
    faabric::Message &A = m[0], &B1 = m[1], &B2 = m[2], &C1 = m[3], &C2 = m[4], &C3 = m[5], &D = m[6];
    faabric::HostResources res;
    res.set_slots(nMsg);
    res.set_usedslots(nMsg);
    f.sch.setThisHostResources(res);
    logChainedFunction(A, B1);
    logChainedFunction(A, B2);
    logChainedFunction(B1, C1);
    logChainedFunction(B2, C2);
    logChainedFunction(B2, C3);
    logChainedFunction(C2, D);
    for (auto& msg : m) {
        f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(msg));
    }
    // (results travel asynchronously)
    for (auto& msg : m) {
        REQUIRE_EQ(f.plannerCli.getMessageResult(msg, 2000).id(), msg.id());
    }
    ExecGraph actual = getFunctionExecGraph(A);
    ExecGraphNode nD{ D, {} }, nC3{ C3, {} }, nC1{ C1, {} };
    ExecGraphNode nC2{ C2, { nD } };
    ExecGraphNode nB2{ B2, { nC2, nC3 } };
    ExecGraphNode nB1{ B1, { nC1 } };
    ExecGraph expected{ ExecGraphNode{ A, { nB1, nB2 } } };
    REQUIRE_EQ(countExecGraphNodes(actual), 7);
    REQUIRE_EQ(countExecGraphNodes(expected), 7);
    REQUIRE(sameGraph(actual.rootNode, expected.rootNode));
}

TEST_CASE("exec graph case: no graph for a call whose result is not published", "[util][exec-graph][cases]")
{
    ClusterFixture f(2);
    faabric::Message msg = messageFactory("demo", "echo");
    REQUIRE_EQ(getFunctionExecGraph(msg).rootNode.msg.id(), 0);
}

TEST_CASE("exec graph case: the set of hosts a graph ran on", "[util][exec-graph][cases]")
{
    auto ber = batchExecFactory("demo", "echo", 3);
    faabric::Message A = ber->messages(0), B1 = ber->messages(1), B2 = ber->messages(2);
    A.set_executedhost("foo");
    B1.set_executedhost("bar");
    B2.set_executedhost("baz");
    ExecGraph graph{ ExecGraphNode{ A, { ExecGraphNode{ B1, {} }, ExecGraphNode{ B2, {} }, ExecGraphNode{ B2, {} } } } };
    REQUIRE(getExecGraphHosts(graph) == (std::set<std::string>{ "bar", "baz", "foo" }));
}

TEST_CASE("exec graph case: details are only recorded once recording is on", "[util][exec-graph][cases]")
{
    faabric::Message msg = messageFactory("foo", "bar");
    REQUIRE(!msg.recordexecgraph());
    incrementCounter(msg, "foo", 1);
    addDetail(msg, "foo", "bar");
    REQUIRE_EQ(msg.intexecgraphdetails_size(), 0);
    REQUIRE_EQ(msg.execgraphdetails_size(), 0);
    msg.set_recordexecgraph(true);
    incrementCounter(msg, "foo", 1);
    addDetail(msg, "foo", "bar");
    REQUIRE_EQ(msg.intexecgraphdetails_size(), 1);
    REQUIRE_EQ(msg.execgraphdetails_size(), 1);
    REQUIRE_EQ(msg.intexecgraphdetails().count("foo"), 1u);
    REQUIRE_EQ(msg.intexecgraphdetails().at("foo"), 1);
    REQUIRE_EQ(msg.execgraphdetails().count("foo"), 1u);
    REQUIRE_EQ(msg.execgraphdetails().at("foo"), std::string("bar"));
}

// ---- concurrent map --------------------------------------------------------------
TEST_CASE("concurrent map case: every operation from one thread", "[util][concurrent_map][cases]")
{
    typedef std::vector<std::pair<int, int>> Pairs;
    const size_t initialCapacity = 32;
    ConcurrentMap<int, int> map(initialCapacity);
    REQUIRE(map.isEmpty());
    REQUIRE_EQ(map.size(), 0u);
    REQUIRE(map.capacity() >= initialCapacity);
    REQUIRE(map.sortedKvPairs().empty());
    map.reserve(2 * initialCapacity);
    REQUIRE(map.capacity() >= 2 * initialCapacity);

    REQUIRE(map.insert(std::make_pair(1, 10)));
    REQUIRE(!map.insert(std::make_pair(1, 20)));
    REQUIRE(map.insert(std::make_pair(3, 30)));
    REQUIRE(!map.isEmpty());
    REQUIRE_EQ(map.size(), 2u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 10 }, { 3, 30 } }));
    map.rehash(0);
    REQUIRE_EQ(map.size(), 2u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 10 }, { 3, 30 } }));

    REQUIRE(!map.insertOrAssign(1, 20));
    REQUIRE(map.insertOrAssign(2, 20));
    REQUIRE_EQ(map.size(), 3u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 20 }, { 2, 20 }, { 3, 30 } }));

    REQUIRE(map.tryEmplace(4, 40));
    REQUIRE(map.tryEmplace(5, 50));
    REQUIRE(!map.tryEmplace(3, 50));
    REQUIRE_EQ(map.size(), 5u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 20 }, { 2, 20 }, { 3, 30 }, { 4, 40 }, { 5, 50 } }));

    map.erase(1);
    REQUIRE_EQ(map.size(), 4u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 2, 20 }, { 3, 30 }, { 4, 40 }, { 5, 50 } }));

    int called = 0;
    bool placedSeen = false;
    int valueSeen = -1;
    map.tryEmplaceThenMutate(
      1,
      [&](bool placed, int& val) {
          called++;
          placedSeen = placed;
          valueSeen = val;
          val = 10;
      },
      0);
    REQUIRE_EQ(called, 1);
    REQUIRE(placedSeen);
    REQUIRE_EQ(valueSeen, 0);
    REQUIRE_EQ(map.size(), 5u);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 10 }, { 2, 20 }, { 3, 30 }, { 4, 40 }, { 5, 50 } }));

    called = 0;
    REQUIRE(map.mutate(5, [&](int& value) {
        called++;
        valueSeen = value;
        value = 51;
    }));
    REQUIRE_EQ(called, 1);
    REQUIRE_EQ(valueSeen, 50);
    called = 0;
    REQUIRE(!map.mutate(7, [&](int& value) {
        called++;
        value = 70;
    }));
    REQUIRE_EQ(called, 0);
    REQUIRE(map.inspect(4, [&](const int& value) {
        called++;
        valueSeen = value;
    }));
    REQUIRE_EQ(called, 1);
    REQUIRE_EQ(valueSeen, 40);
    called = 0;
    REQUIRE(map.inspect(4, [&](int value) { called++; }));
    REQUIRE_EQ(called, 1);
    called = 0;
    REQUIRE(!map.inspect(7, [&](const int&) { called++; }));
    REQUIRE_EQ(called, 0);
    REQUIRE(map.get(7) == std::nullopt);
    REQUIRE(!map.contains(7));
    REQUIRE(map.get(4) == 40);
    REQUIRE(map.contains(4));
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 10 }, { 2, 20 }, { 3, 30 }, { 4, 40 }, { 5, 51 } }));

    called = 0;
    int sum = 0;
    map.inspectAll([&](const int&, const int& value) {
        called++;
        sum += value;
    });
    REQUIRE_EQ(sum, 10 + 20 + 30 + 40 + 51);
    REQUIRE_EQ(called, 5);
    called = 0;
    map.mutateAll([&](const int&, int& value) {
        called++;
        value /= 10;
    });
    REQUIRE_EQ(called, 5);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 1, 1 }, { 2, 2 }, { 3, 3 }, { 4, 4 }, { 5, 5 } }));
    map.eraseIf([](const int& key, const int&) { return key <= 3; });
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 4, 4 }, { 5, 5 } }));

    // swapping exchanges contents
    ConcurrentMap<int, int> other;
    other.insert(std::make_pair(9, 90));
    map.swap(other);
    REQUIRE(map.sortedKvPairs() == (Pairs{ { 9, 90 } }));
    REQUIRE(other.sortedKvPairs() == (Pairs{ { 4, 4 }, { 5, 5 } }));

    map.clear();
    REQUIRE(map.isEmpty());
    REQUIRE_EQ(map.size(), 0u);
    REQUIRE(map.sortedKvPairs().empty());
}

TEST_CASE("concurrent map case: insertion from many threads, through a resize", "[util][concurrent_map][cases]")
{
    const int nThreads = 5, nValues = 1000;
    ConcurrentMap<int, int> map(nValues); // undersized on purpose
    std::vector<std::thread> workers;
    for (int t = 0; t < nThreads; t++) {
        workers.emplace_back([t, &map] {
            for (int i = 0; i < nValues; i++) {
                map.tryEmplace(nThreads * i + t, i);
            }
        });
    }
    for (auto& w : workers) {
        w.join();
    }
    auto values = map.sortedKvPairs();
    REQUIRE_EQ(values.size(), (size_t)nThreads * nValues);
    bool ok = true;
    for (int i = 0; i < nValues; i++) {
        for (int t = 0; t < nThreads; t++) {
            int key = nThreads * i + t;
            ok = ok && values.at(key).first == key && values.at(key).second == i;
        }
    }
    REQUIRE(ok);
}

// ---- periodic thread ---------------------------------------------------------------
namespace {
class CountingPeriodicThread : public PeriodicBackgroundThread
{
  public:
    explicit CountingPeriodicThread(std::shared_ptr<Barrier> b)
      : barrier(std::move(b))
    {}
    void doWork() override
    {
        isRunning = true;
        workCount++;
        barrier->wait();
    }
    void tidyUp() override { isRunning = false; }
    std::atomic<bool> isRunning{ false };
    std::atomic<int> workCount{ 0 };

  private:
    std::shared_ptr<Barrier> barrier;
};
}

TEST_CASE("periodic thread case: works once per interval until stopped, then tidies up", "[util][periodic][cases]")
{
    auto b = Barrier::create(2);
    CountingPeriodicThread t(b);
    REQUIRE_EQ(t.workCount.load(), 0);
    t.start(1);
    b->wait();
    REQUIRE_EQ(t.workCount.load(), 1);
    REQUIRE(t.isRunning.load());
    b->wait();
    REQUIRE_EQ(t.workCount.load(), 2);
    t.stop();
    REQUIRE(!t.isRunning.load());
    REQUIRE_EQ(t.workCount.load(), 2);
}

TEST_CASE("periodic thread case: a non-positive interval never starts it", "[util][periodic][cases]")
{
    auto b = Barrier::create(2);
    CountingPeriodicThread t(b);
    t.start(0);
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    REQUIRE(!t.isRunning.load());
    REQUIRE_EQ(t.workCount.load(), 0);
    t.start(-3);
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
    REQUIRE(!t.isRunning.load());
    t.stop(); // stopping what never started is fine
}

// ---- messages ------------------------------------------------------------------------
TEST_CASE("proto case: a message's fields survive serialisation", "[proto][cases]")
{
    faabric::Message call;
    std::vector<uint8_t> input;
    for (int i = 0; i < 100; i++) {
        input.push_back((uint8_t)i);
    }
    call.set_user("foobar user");
    call.set_function("foobar func");
    call.set_resultkey("dummy result");
    call.set_returnvalue(123);
    call.set_pythonuser("python user");
    call.set_pythonfunction("python func");
    call.set_pythonentry("python entry");
    call.set_inputdata(input.data(), 100);
    call.set_outputdata("output data");
    call.set_ispython(true);
    call.set_type(faabric::Message_MessageType_KILL);
    call.set_cmdline("some cmdline args");
    REQUIRE(call.type() == faabric::Message_MessageType_KILL);
    REQUIRE_EQ(call.returnvalue(), 123);

    std::string wire = call.SerializeAsString();
    faabric::Message back;
    REQUIRE(back.ParseFromString(wire));
    REQUIRE_EQ(back.user(), std::string("foobar user"));
    REQUIRE_EQ(back.function(), std::string("foobar func"));
    REQUIRE_EQ(back.resultkey(), std::string("dummy result"));
    REQUIRE(back.type() == faabric::Message_MessageType_KILL);
    REQUIRE_EQ(back.pythonuser(), std::string("python user"));
    REQUIRE_EQ(back.pythonfunction(), std::string("python func"));
    REQUIRE_EQ(back.pythonentry(), std::string("python entry"));
    REQUIRE(back.ispython());
    REQUIRE_EQ(back.cmdline(), std::string("some cmdline args"));
    REQUIRE(stringToBytes(back.inputdata()) == input);
    REQUIRE_EQ(back.outputdata(), std::string("output data"));
}

TEST_CASE("proto case: byte fields keep embedded zeros", "[proto][cases]")
{
    faabric::Message a, b;
    std::vector<uint8_t> bytesA = { 0, 0, 1, 1, 0, 0, 2, 2 };
    std::vector<uint8_t> bytesB = { 1, 1, 1, 1, 1, 1, 2, 2 };
    a.set_inputdata(bytesA.data(), bytesA.size());
    b.set_inputdata(bytesB.data(), bytesB.size());
    std::string wireA = a.SerializeAsString(), wireB = b.SerializeAsString();
    REQUIRE_EQ(wireA.size(), wireB.size());
    faabric::Message backA, backB;
    REQUIRE(backA.ParseFromString(wireA));
    REQUIRE(backB.ParseFromString(wireB));
    REQUIRE(stringToBytes(backA.inputdata()) == bytesA);
    REQUIRE(stringToBytes(backB.inputdata()) == bytesB);
}

// ---- dirty trackers ---------------------------------------------------------------------
namespace {
struct TrackerScope
{
    explicit TrackerScope(const std::string& mode)
    {
        getSystemConfig().dirtyTrackingMode = mode;
        resetDirtyTracker();
    }
    ~TrackerScope()
    {
        getSystemConfig().reset();
        resetDirtyTracker();
    }
};

bool trackerUsable(const std::string& mode)
{
    if (mode == "softpte") {
        return SoftPTEDirtyTracker::isSupported();
    }
    if (mode.rfind("uffd", 0) == 0) {
        return UffdDirtyTracker::isSupported();
    }
    return true;
}

enum MemKind
{
    PRIVATE_MEM,
    SHARED_MEM,
    MAPPED_PRIVATE,
    MAPPED_SHARED
};

// The reference's "basic dirty tracking" body for one (tracker, memory) pair
void basicTracking(const std::string& mode, MemKind kind, bool checkPostReset)
{
    TrackerScope scope(mode);
    const int nPages = 6;
    const size_t memSize = (size_t)HOST_PAGE_SIZE * nPages;
    MemoryRegion sharedMem = allocateSharedMemory(memSize);
    MemoryRegion privateMem = allocatePrivateMemory(memSize);
    bool shared = kind == SHARED_MEM || kind == MAPPED_SHARED;
    std::span<uint8_t> mem(shared ? sharedMem.get() : privateMem.get(), memSize);
    int fd = -1;
    if (kind == MAPPED_PRIVATE || kind == MAPPED_SHARED) {
        fd = createFd(memSize, "foobar");
        if (shared) {
            mapMemoryShared(mem, fd);
        } else {
            mapMemoryPrivate(mem, fd);
        }
    }
    auto tracker = getDirtyTracker();
    REQUIRE_EQ(tracker->getType(), mode);
    tracker->clearAll();
    std::vector<char> expected(nPages, 0);
    REQUIRE(tracker->getBothDirtyPages(mem) == expected);

    tracker->startTracking(mem);
    tracker->startThreadLocalTracking(mem);
    uint8_t* page0 = mem.data();
    uint8_t* page1 = page0 + HOST_PAGE_SIZE;
    uint8_t* page3 = page1 + 2 * HOST_PAGE_SIZE;
    uint8_t* page4 = page3 + HOST_PAGE_SIZE;
    uint8_t* page5 = page3 + 2 * HOST_PAGE_SIZE;
    // a read marks nothing (every tracker here works on write faults)
    volatile int readValue = page0[1];
    REQUIRE_EQ((int)readValue, 0);
    page1[10] = 1;
    page3[123] = 4;
    expected = { 0, 1, 0, 1, 0, 0 };
    REQUIRE(tracker->getBothDirtyPages(mem) == expected);
    page5[99] = 3;
    expected[5] = 1;
    REQUIRE(tracker->getBothDirtyPages(mem) == expected);

    // stop and start again: a clean slate, data untouched
    tracker->stopTracking(mem);
    tracker->stopThreadLocalTracking(mem);
    tracker->startTracking(mem);
    tracker->startThreadLocalTracking(mem);
    REQUIRE(tracker->getBothDirtyPages(mem) == std::vector<char>(nPages, 0));
    REQUIRE_EQ((int)page1[10], 1);
    REQUIRE_EQ((int)page3[123], 4);
    REQUIRE_EQ((int)page5[99], 3);

    if (checkPostReset) {
        page3[100] = 2;
        page4[22] = 5;
        expected = std::vector<char>(nPages, 0);
        expected[3] = expected[4] = 1;
        REQUIRE(tracker->getBothDirtyPages(mem) == expected);
        tracker->stopTracking(mem);
        tracker->stopThreadLocalTracking(mem);
        tracker->startTracking(mem);
        tracker->startThreadLocalTracking(mem);
        REQUIRE(tracker->getBothDirtyPages(mem) == std::vector<char>(nPages, 0));
    }
    tracker->stopTracking(mem);
    tracker->stopThreadLocalTracking(mem);
    // writable again
    page0[0] = 9;
    if (fd >= 0) {
        ::close(fd);
    }
}

// The reference's thread-local body: 100 threads, two pages each, repeated
void threadLocalTracking(const std::string& mode, int nLoops)
{
    TrackerScope scope(mode);
    auto tracker = getDirtyTracker();
    REQUIRE_EQ(tracker->getType(), mode);
    const int nThreads = 100;
    const int nPages = 2 * nThreads;
    const size_t memSize = (size_t)nPages * HOST_PAGE_SIZE;
    MemoryRegion region = allocatePrivateMemory(memSize);
    std::span<uint8_t> mem(region.get(), memSize);
    for (int loop = 0; loop < nLoops; loop++) {
        std::vector<std::atomic<int>> outcome(nThreads);
        tracker->startTracking(mem);
        std::vector<std::thread> threads;
        for (int i = 0; i < nThreads; i++) {
            threads.emplace_back([&, i] {
                tracker->startThreadLocalTracking(mem);
                size_t pageOffset = (size_t)i * 2;
                uint8_t* one = mem.data() + pageOffset * HOST_PAGE_SIZE;
                uint8_t* two = one + HOST_PAGE_SIZE;
                one[20] = 3;
                one[250] = 5;
                one[HOST_PAGE_SIZE - 20] = 6;
                two[35] = 2;
                two[HOST_PAGE_SIZE - 100] = 3;
                tracker->stopThreadLocalTracking(mem);
                auto dirty = tracker->getThreadLocalDirtyPages(mem);
                std::vector<char> expected(nPages, 0);
                expected[pageOffset] = expected[pageOffset + 1] = 1;
                outcome[i] = dirty == expected ? 1 : -1;
            });
        }
        for (auto& t : threads) {
            t.join();
        }
        tracker->stopTracking(mem);
        // nothing was written outside thread-local tracking
        auto global = tracker->getDirtyPages(mem);
        REQUIRE_EQ(std::count(global.begin(), global.end(), 1), 0);
        int failed = 0;
        for (auto& o : outcome) {
            failed += o.load() == 1 ? 0 : 1;
        }
        REQUIRE_EQ(failed, 0);
    }
}
}

TEST_CASE("dirty case: the configured mode names the tracker, all seven of them", "[util][dirty][cases]")
{
    for (const char* mode : { "segfault", "softpte", "none", "uffd", "uffd-wp", "uffd-thread", "uffd-thread-wp" }) {
        if (!trackerUsable(mode)) {
            continue;
        }
        TrackerScope scope(mode);
        REQUIRE_EQ(getDirtyTracker()->getType(), std::string(mode));
    }
}

#define DIRTY_CASE(label, mode, kind, postReset)                                                                       \
    TEST_CASE("dirty case: " label, "[util][dirty][cases]")                                                            \
    {                                                                                                                  \
        if (!trackerUsable(mode)) {                                                                                    \
            SKIP_TEST(mode " tracking is not available here");                                                         \
        }                                                                                                              \
        basicTracking(mode, kind, postReset);                                                                          \
    }

DIRTY_CASE("segfault tracker, private memory", "segfault", PRIVATE_MEM, true)
DIRTY_CASE("segfault tracker, shared memory", "segfault", SHARED_MEM, true)
DIRTY_CASE("segfault tracker, file-mapped private memory", "segfault", MAPPED_PRIVATE, true)
DIRTY_CASE("segfault tracker, file-mapped shared memory", "segfault", MAPPED_SHARED, true)
DIRTY_CASE("soft-dirty tracker, private memory", "softpte", PRIVATE_MEM, true)
DIRTY_CASE("soft-dirty tracker, shared memory", "softpte", SHARED_MEM, true)
DIRTY_CASE("soft-dirty tracker, file-mapped private memory", "softpte", MAPPED_PRIVATE, true)
DIRTY_CASE("soft-dirty tracker, file-mapped shared memory", "softpte", MAPPED_SHARED, true)
DIRTY_CASE("uffd tracker, private memory", "uffd", PRIVATE_MEM, true)
DIRTY_CASE("uffd tracker, shared memory", "uffd", SHARED_MEM, true)
DIRTY_CASE("uffd tracker, file-mapped private memory", "uffd", MAPPED_PRIVATE, true)
DIRTY_CASE("uffd tracker, file-mapped shared memory", "uffd", MAPPED_SHARED, true)
DIRTY_CASE("uffd-wp tracker, private memory", "uffd-wp", PRIVATE_MEM, true)
DIRTY_CASE("uffd-thread tracker, private memory", "uffd-thread", PRIVATE_MEM, true)
DIRTY_CASE("uffd-thread tracker, shared memory", "uffd-thread", SHARED_MEM, true)
DIRTY_CASE("uffd-thread tracker, file-mapped private memory", "uffd-thread", MAPPED_PRIVATE, true)
DIRTY_CASE("uffd-thread tracker, file-mapped shared memory", "uffd-thread", MAPPED_SHARED, true)
DIRTY_CASE("uffd-thread-wp tracker, private memory", "uffd-thread-wp", PRIVATE_MEM, true)

TEST_CASE("dirty case: uffd tracker with only the basic kernel features pre-faults untouched pages", "[util][dirty][cases]")
{
    if (!trackerUsable("uffd")) {
        SKIP_TEST("userfaultfd write-protect is not available here");
    }
    ::setenv("FAABRIC_UFFD_FEATURES", "basic", 1);
    std::shared_ptr<void> restore(nullptr, [](void*) { ::unsetenv("FAABRIC_UFFD_FEATURES"); });
    basicTracking("uffd", PRIVATE_MEM, true);
}

TEST_CASE("dirty case: a hundred threads each see only their own pages (segfault tracker, 20 rounds)", "[util][dirty][cases]")
{
    threadLocalTracking("segfault", 20);
}

TEST_CASE("dirty case: a hundred threads each see only their own pages (uffd-wp tracker)", "[util][dirty][cases]")
{
    if (!trackerUsable("uffd-wp")) {
        SKIP_TEST("userfaultfd write-protect is not available here");
    }
    threadLocalTracking("uffd-wp", 20);
}
