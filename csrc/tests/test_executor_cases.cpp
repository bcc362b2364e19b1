// Executor behaviour, one case per scenario of the reference's executor suite
// (tests/test/executor/test_executor.cpp: "Test executing simple function",
// "... repeatedly and flushing", "... chained functions", "... threads
// directly", "... chained threads", "Test non-zero return code", "Test erroring
// function / thread", "Test executing different functions", "Test snapshot
// diffs returned to main", "Test single host flag passed to executor", "Test
// executor sees context", "Test executor restore", "Test get main thread
// snapshot", "Test executor keeps track of chained messages", "Test executing
// threads manually").  Scenarios are re-implemented against this repo's
// fixtures; nothing is taken from the reference's test code.
#include "fixtures.h"

#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/util/memory.h>
#include <faabric/util/snapshot.h>

#include <thread>

using namespace tests;
using faabric::util::SnapshotDataType;
using faabric::util::SnapshotMergeOperation;

namespace {
std::shared_ptr<faabric::BatchExecuteRequest> threadsOf(const faabric::Message& parent,
                                                        const std::string& function,
                                                        int n)
{
    auto threads = faabric::util::batchExecFactory(parent.user(), function, n);
    threads->set_type(faabric::BatchExecuteRequest::THREADS);
    faabric::util::updateBatchExecAppId(threads, parent.appid());
    for (int i = 0; i < n; i++) {
        threads->mutable_messages(i)->set_appidx(i + 1);
        threads->mutable_messages(i)->set_groupidx(i + 1);
    }
    return threads;
}
}

TEST_CASE("executor case: a simple function runs and returns its output", "[executor][cases]")
{
    ClusterFixture f(4);
    registerTestFunction("cases", "simple", [](auto*, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        m.set_outputdata("Simple function " + std::to_string(m.id()) + " executed");
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "simple", 1);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.hosts.size(), 1u);
    auto res = f.awaitResult(req->messages(0));
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), "Simple function " + std::to_string(req->messages(0).id()) + " executed");
    REQUIRE_EQ(res.executedhost(), f.conf.endpointHost);
    REQUIRE(res.finishtimestamp() >= res.starttimestamp());
}

TEST_CASE("executor case: repeated execution reuses the warm executor until a flush", "[executor][cases]")
{
    ClusterFixture f(4);
    std::atomic<int> runs{ 0 };
    registerTestFunction("cases", "again", [&](auto*, int, int, auto) {
        runs++;
        return 0;
    });
    faabric::Message first;
    for (int i = 0; i < 5; i++) {
        auto req = faabric::util::batchExecFactory("cases", "again", 1);
        if (i == 0) {
            first = req->messages(0);
        }
        f.plannerCli.callFunctions(req);
        REQUIRE_EQ(f.awaitResult(req->messages(0)).returnvalue(), 0);
        f.awaitBatch(req);
    }
    REQUIRE_EQ(runs.load(), 5);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(first), 1);
    // a flush drops the executors and tells the factory
    faabric::scheduler::getFunctionCallClient(f.conf.endpointHost)->sendFlush();
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(first), 0);
    REQUIRE_EQ(f.factory->flushCount, 1);
    auto req = faabric::util::batchExecFactory("cases", "again", 1);
    f.plannerCli.callFunctions(req);
    REQUIRE_EQ(f.awaitResult(req->messages(0)).returnvalue(), 0);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(first), 1);
}

TEST_CASE("executor case: chained functions are invoked from inside a function", "[executor][cases]")
{
    ClusterFixture f(8);
    registerTestFunction("cases", "chain-child", [](auto*, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        m.set_outputdata("child got " + m.inputdata());
        return 0;
    });
    registerTestFunction("cases", "chain-parent", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto children = faabric::util::batchExecFactory("cases", "chain-child", 3);
        faabric::util::updateBatchExecAppId(children, m.appid());
        for (int i = 0; i < 3; i++) {
            children->mutable_messages(i)->set_inputdata("chained-msg-" + std::to_string(i));
            children->mutable_messages(i)->set_appidx(i + 1);
            exec->addChainedMessage(children->messages(i));
        }
        faabric::planner::getPlannerClient().callFunctions(children);
        std::string all;
        for (int i = 0; i < 3; i++) {
            auto r = faabric::planner::getPlannerClient().getMessageResult(children->messages(i), 10000);
            if (r.returnvalue() != 0) {
                return 1;
            }
            all += r.outputdata() + ";";
            // the parent can look its children up again
            if (exec->getChainedMessage(children->messages(i).id()).inputdata() != "chained-msg-" + std::to_string(i)) {
                return 2;
            }
        }
        m.set_outputdata(all);
        return (int)exec->getChainedMessageIds().size() == 3 ? 0 : 3;
    });
    auto req = faabric::util::batchExecFactory("cases", "chain-parent", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0), 20000);
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::string("child got chained-msg-0;child got chained-msg-1;child got chained-msg-2;"));
}

TEST_CASE("executor case: unknown chained message ids are reported", "[executor][cases]")
{
    ClusterFixture f(2);
    auto msg = faabric::util::messageFactory("cases", "tracker");
    TestExecutor exec(msg);
    auto a = faabric::util::messageFactory("cases", "a");
    auto b = faabric::util::messageFactory("cases", "b");
    exec.addChainedMessage(a);
    exec.addChainedMessage(b);
    REQUIRE_EQ(exec.getChainedMessageIds().size(), 2u);
    REQUIRE_EQ(exec.getChainedMessage(a.id()).function(), std::string("a"));
    REQUIRE_THROWS(exec.getChainedMessage(123456789));
    exec.shutdown();
}

TEST_CASE("executor case: threads executed directly return one result per thread", "[executor][cases]")
{
    ClusterFixture f(8);
    const int nThreads = 6;
    registerTestFunction("cases", "thread-body", [](auto*, int, int idx, auto req) { return req->messages(idx).appidx() * 10; });
    registerTestFunction("cases", "thread-main", [&](auto* exec, int, int idx, auto req) {
        auto threads = threadsOf(req->messages(idx), "thread-body", nThreads);
        threads->set_singlehost(true);
        auto results = exec->executeThreads(threads, {});
        if ((int)results.size() != nThreads) {
            return 1;
        }
        std::map<uint32_t, int32_t> byId(results.begin(), results.end());
        for (int i = 0; i < nThreads; i++) {
            if (byId[threads->messages(i).id()] != (i + 1) * 10) {
                return 2;
            }
        }
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "thread-main", 1);
    f.plannerCli.callFunctions(req);
    REQUIRE_EQ(f.awaitResult(req->messages(0), 20000).returnvalue(), 0);
}

TEST_CASE("executor case: threads forked repeatedly from the same function", "[executor][cases]")
{
    ClusterFixture f(6);
    std::atomic<int> bodies{ 0 };
    registerTestFunction("cases", "rep-body", [&](auto*, int, int, auto) {
        bodies++;
        return 0;
    });
    registerTestFunction("cases", "rep-main", [&](auto* exec, int, int idx, auto req) {
        for (int round = 0; round < 5; round++) {
            auto threads = threadsOf(req->messages(idx), "rep-body", 3);
            threads->set_singlehost(true);
            if (exec->executeThreads(threads, {}).size() != 3) {
                return 1;
            }
        }
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "rep-main", 1);
    f.plannerCli.callFunctions(req);
    REQUIRE_EQ(f.awaitResult(req->messages(0), 30000).returnvalue(), 0);
    REQUIRE_EQ(bodies.load(), 15);
    // everything ran in the main function's executor
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 1);
}

TEST_CASE("executor case: a non-zero return code reaches the caller", "[executor][cases]")
{
    ClusterFixture f(2);
    registerTestFunction("cases", "ret-one", [](auto*, int, int, auto) { return 1; });
    registerTestFunction("cases", "ret-big", [](auto*, int, int, auto) { return 4242; });
    auto a = faabric::util::batchExecFactory("cases", "ret-one", 1);
    auto b = faabric::util::batchExecFactory("cases", "ret-big", 1);
    f.plannerCli.callFunctions(a);
    f.plannerCli.callFunctions(b);
    REQUIRE_EQ(f.awaitResult(a->messages(0)).returnvalue(), 1);
    REQUIRE_EQ(f.awaitResult(b->messages(0)).returnvalue(), 4242);
}

TEST_CASE("executor case: an exception in a function becomes return code 1 plus a message", "[executor][cases]")
{
    ClusterFixture f(2);
    registerTestFunction("cases", "thrower", [](auto*, int, int, auto) -> int { throw std::runtime_error("this function is broken"); });
    auto req = faabric::util::batchExecFactory("cases", "thrower", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0));
    REQUIRE_EQ(res.returnvalue(), 1);
    REQUIRE(res.outputdata().find("this function is broken") != std::string::npos);
    REQUIRE(res.outputdata().find(std::to_string(req->messages(0).id())) != std::string::npos);
    // the executor is usable afterwards
    registerTestFunction("cases", "thrower", [](auto*, int, int, auto) { return 0; });
    auto again = faabric::util::batchExecFactory("cases", "thrower", 1);
    f.plannerCli.callFunctions(again);
    REQUIRE_EQ(f.awaitResult(again->messages(0)).returnvalue(), 0);
}

TEST_CASE("executor case: an exception in one thread does not take the others down", "[executor][cases]")
{
    ClusterFixture f(6);
    registerTestFunction("cases", "thread-err-body", [](auto*, int, int idx, auto req) -> int {
        if (req->messages(idx).appidx() == 2) {
            throw std::runtime_error("thread 2 fails");
        }
        return 0;
    });
    registerTestFunction("cases", "thread-err-main", [&](auto* exec, int, int idx, auto req) {
        auto threads = threadsOf(req->messages(idx), "thread-err-body", 4);
        threads->set_singlehost(true);
        auto results = exec->executeThreads(threads, {});
        int failed = 0;
        for (auto& [id, rv] : results) {
            failed += rv != 0;
        }
        return results.size() == 4 && failed == 1 ? 0 : 1;
    });
    auto req = faabric::util::batchExecFactory("cases", "thread-err-main", 1);
    f.plannerCli.callFunctions(req);
    REQUIRE_EQ(f.awaitResult(req->messages(0), 20000).returnvalue(), 0);
}

TEST_CASE("executor case: different functions get executors of their own", "[executor][cases]")
{
    ClusterFixture f(8);
    registerTestFunction("cases", "kind-a", [](auto*, int, int idx, auto req) {
        req->mutable_messages(idx)->set_outputdata("A");
        return 0;
    });
    registerTestFunction("cases", "kind-b", [](auto*, int, int idx, auto req) {
        req->mutable_messages(idx)->set_outputdata("B");
        return 0;
    });
    auto a = faabric::util::batchExecFactory("cases", "kind-a", 2);
    auto b = faabric::util::batchExecFactory("cases", "kind-b", 3);
    f.plannerCli.callFunctions(a);
    f.plannerCli.callFunctions(b);
    for (int i = 0; i < 2; i++) {
        REQUIRE_EQ(f.awaitResult(a->messages(i)).outputdata(), std::string("A"));
    }
    for (int i = 0; i < 3; i++) {
        REQUIRE_EQ(f.awaitResult(b->messages(i)).outputdata(), std::string("B"));
    }
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(a->messages(0)), 2);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(b->messages(0)), 3);
}

TEST_CASE("executor case: claiming is exclusive until released", "[executor][cases]")
{
    ClusterFixture f(2);
    auto msg = faabric::util::messageFactory("cases", "claim");
    auto exec = std::make_shared<TestExecutor>(msg);
    REQUIRE(!exec->isExecuting());
    REQUIRE(exec->tryClaim());
    REQUIRE(exec->isExecuting());
    REQUIRE(!exec->tryClaim());
    REQUIRE_THROWS(exec->claim());
    exec->releaseClaim();
    REQUIRE(!exec->isExecuting());
    exec->claim();
    REQUIRE(exec->isExecuting());
    exec->releaseClaim();
    exec->shutdown();
}

TEST_CASE("executor case: the single-host flag of the batch is what the tasks see", "[executor][cases]")
{
    ClusterFixture f(6);
    std::atomic<int> sawSingle{ 0 };
    std::atomic<int> sawSpread{ 0 };
    registerTestFunction("cases", "flag-body", [&](auto*, int, int, auto req) {
        if (req->singlehost()) {
            sawSingle++;
        } else {
            sawSpread++;
        }
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "flag-body", 3);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    // one host in the cluster: the planner marks the dispatched batch as single-host
    REQUIRE_EQ(sawSingle.load(), 3);
    REQUIRE_EQ(sawSpread.load(), 0);
}

TEST_CASE("executor case: the context names executor, request and message index", "[executor][cases]")
{
    ClusterFixture f(4);
    std::atomic<int> mismatches{ 0 };
    REQUIRE(!faabric::executor::ExecutorContext::isSet());
    REQUIRE_THROWS(faabric::executor::ExecutorContext::get());
    registerTestFunction("cases", "ctx", [&](auto* exec, int, int idx, auto req) {
        auto ctx = faabric::executor::ExecutorContext::get();
        if (ctx->getExecutor() != exec || ctx->getBatchRequest() != req || ctx->getMsgIdx() != idx ||
            ctx->getMsg().id() != req->messages(idx).id()) {
            mismatches++;
        }
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "ctx", 3);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    REQUIRE_EQ(mismatches.load(), 0);
    REQUIRE(!faabric::executor::ExecutorContext::isSet());
}

TEST_CASE("executor case: restore maps the registered image into the executor's memory", "[executor][cases]")
{
    ClusterFixture f(2);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    const size_t size = 4 * faabric::util::HOST_PAGE_SIZE;
    std::vector<uint8_t> image(size);
    for (size_t i = 0; i < size; i++) {
        image[i] = (uint8_t)(i % 199);
    }
    reg.registerSnapshot("cases-restore", std::make_shared<faabric::util::SnapshotData>(std::span<const uint8_t>(image.data(), size)));
    auto msg = faabric::util::messageFactory("cases", "restore");
    TestExecutor exec(msg);
    int before = TestExecutor::restoreCount.load();
    exec.restore("cases-restore");
    REQUIRE_EQ(TestExecutor::restoreCount.load(), before + 1);
    auto view = exec.getMemoryView();
    REQUIRE_EQ(view.size(), size);
    REQUIRE(memcmp(view.data(), image.data(), size) == 0);
    // the mapping is private to the executor
    view[17] = 0xab;
    REQUIRE_EQ((int)*reg.getSnapshot("cases-restore")->getDataPtr(17), (int)image[17]);
    REQUIRE_THROWS(exec.restore("no-such-snapshot"));
    exec.shutdown();
}

TEST_CASE("executor case: the main thread snapshot is created on demand and keyed by the app", "[executor][cases]")
{
    ClusterFixture f(2);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    auto msg = faabric::util::messageFactory("cases", "main-snap");
    msg.set_appid(991122);
    TestExecutor exec(msg);
    std::string key = faabric::util::getMainThreadSnapshotKey(msg);
    REQUIRE(!reg.snapshotExists(key));
    REQUIRE_THROWS(exec.getMainThreadSnapshot(msg, false));
    exec.getMemoryView()[40] = 0x7c;
    auto snap = exec.getMainThreadSnapshot(msg, true);
    REQUIRE(reg.snapshotExists(key));
    REQUIRE_EQ(snap->getSize(), exec.getMemoryView().size());
    REQUIRE_EQ((int)*snap->getDataPtr(40), 0x7c);
    // asking again returns the same object
    REQUIRE(exec.getMainThreadSnapshot(msg, true) == snap);
    // another app of the same function has its own key
    auto other = msg;
    other.set_appid(991123);
    REQUIRE(faabric::util::getMainThreadSnapshotKey(other) != key);
    exec.shutdown();
}

TEST_CASE("executor case: thread diffs of a remote host travel back to the main host", "[executor][cases]")
{
    // The executor plays the role of a NON-main host: its threads start from
    // the main thread's snapshot, and what they change goes back as diffs
    ClusterFixture f(4);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    const size_t size = 16 * faabric::util::HOST_PAGE_SIZE;
    std::vector<uint8_t> image(size, 0);
    int counter = 10;
    memcpy(image.data() + 128, &counter, 4);
    auto msg = faabric::util::messageFactory("cases", "remote-threads");
    msg.set_appid(8800);
    std::string key = faabric::util::getMainThreadSnapshotKey(msg);
    auto snap = std::make_shared<faabric::util::SnapshotData>(std::span<const uint8_t>(image.data(), size));
    snap->addMergeRegion(128, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    reg.registerSnapshot(key, snap);

    registerTestFunction("cases", "remote-threads", [](auto* exec, int, int idx, auto req) {
        auto mem = exec->getMemoryView();
        int t = req->messages(idx).appidx();
        mem[2 * faabric::util::HOST_PAGE_SIZE + t] = (uint8_t)(0x40 + t);
        __atomic_fetch_add((int*)(mem.data() + 128), t, __ATOMIC_RELAXED);
        return 0;
    });
    auto req = faabric::util::batchExecFactory("cases", "remote-threads", 3);
    req->set_type(faabric::BatchExecuteRequest::THREADS);
    req->set_singlehost(false);
    faabric::util::updateBatchExecAppId(req, 8800);
    for (int i = 0; i < 3; i++) {
        auto* m = req->mutable_messages(i);
        m->set_appidx(i + 1);
        m->set_groupidx(i + 1);
        m->set_mainhost("the-main-host");
        m->set_executedhost(f.conf.endpointHost);
    }
    faabric::util::setMockMode(true);
    faabric::snapshot::clearMockSnapshotRequests();
    auto exec = std::make_shared<TestExecutor>(*req->mutable_messages(0));
    exec->claim();
    exec->executeTasks({ 0, 1, 2 }, req);
    // mock mode records what would have gone over the wire
    std::vector<std::pair<std::string, std::tuple<int, int, std::string, int>>> results;
    for (int i = 0; i < 400; i++) {
        results = faabric::snapshot::getThreadResults();
        if (results.size() >= 3) {
            break;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    faabric::util::setMockMode(false);
    // every remote thread reports to the main host (reference:
    // src/executor/Executor.cpp:296-300); only the last one of the batch
    // carries the snapshot key and the diffs
    REQUIRE_EQ(results.size(), 3u);
    int withDiffs = 0;
    for (auto& r : results) {
        REQUIRE_EQ(r.first, std::string("the-main-host"));
        if (std::get<3>(r.second) > 0) {
            withDiffs++;
            REQUIRE_EQ(std::get<2>(r.second), key);
            // the Sum region + the bytes of three threads in one page
            REQUIRE(std::get<3>(r.second) >= 2);
        } else {
            REQUIRE(std::get<2>(r.second).empty());
        }
    }
    REQUIRE_EQ(withDiffs, 1);
    exec->shutdown();
    faabric::snapshot::clearMockSnapshotRequests();
}

TEST_CASE("executor case: tasks can be handed to an executor by hand", "[executor][cases]")
{
    ClusterFixture f(4);
    std::atomic<int> ran{ 0 };
    registerTestFunction("cases", "manual", [&](auto*, int, int idx, auto req) {
        ran++;
        return 100 + req->messages(idx).appidx();
    });
    auto req = faabric::util::batchExecFactory("cases", "manual", 3);
    for (int i = 0; i < 3; i++) {
        req->mutable_messages(i)->set_appidx(i);
    }
    // schedule through the planner so that results have somewhere to go
    f.plannerCli.callFunctions(req);
    auto status = f.awaitBatch(req);
    REQUIRE_EQ(ran.load(), 3);
    std::set<int> rvs;
    for (auto& m : status->messageresults()) {
        rvs.insert(m.returnvalue());
    }
    REQUIRE(rvs == (std::set<int>{ 100, 101, 102 }));
}

TEST_CASE("executor case: the time since the last execution restarts with every execution", "[executor][cases]")
{
    ClusterFixture f(5);
    auto req = faabric::util::batchExecFactory("foo", "bar", 1);
    faabric::Message msg = req->messages(0);
    req->mutable_messages(0)->set_executedhost(f.conf.endpointHost);
    faabric::HostResources res;
    res.set_slots(5);
    res.set_usedslots(5);
    f.sch.setThisHostResources(res);
    f.sch.addHostToGlobalSet();
    auto exec = std::make_shared<TestExecutor>(*req->mutable_messages(0));
    long a = exec->getMillisSinceLastExec();
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    long b = exec->getMillisSinceLastExec();
    REQUIRE(b > a);
    REQUIRE(b - a > 90);
    exec->claim();
    exec->executeTasks({ 0 }, req);
    REQUIRE(exec->getMillisSinceLastExec() < b);
    f.awaitResult(msg, 2000);
    exec->shutdown();
}

TEST_CASE("executor case: chained messages are kept until the executor is reset", "[executor][cases]")
{
    ClusterFixture f(2);
    auto req = faabric::util::batchExecFactory("hello", "world", 1);
    auto& first = *req->mutable_messages(0);
    auto exec = std::make_shared<TestExecutor>(first);
    REQUIRE(exec->getChainedMessageIds().empty());
    faabric::Message chained = faabric::util::messageFactory("hello", "chained");
    chained.set_inputdata("payload");
    exec->addChainedMessage(chained);
    REQUIRE(exec->getChainedMessageIds() == (std::set<unsigned int>{ (unsigned int)chained.id() }));
    const faabric::Message& kept = exec->getChainedMessage(chained.id());
    REQUIRE_EQ(kept.id(), chained.id());
    REQUIRE_EQ(kept.function(), std::string("chained"));
    REQUIRE_EQ(kept.inputdata(), std::string("payload"));
    REQUIRE_THROWS(exec->getChainedMessage(chained.id() + 1));
    exec->reset(first);
    REQUIRE(exec->getChainedMessageIds().empty());
    exec->shutdown();
}

TEST_CASE("executor case: executors idle for longer than the bound timeout are reaped, busy ones are not", "[executor][cases]")
{
    ClusterFixture f(4);
    f.conf.boundTimeout = 1500;
    auto release = std::make_shared<std::atomic<bool>>(false);
    // (whatever an assertion below does, the busy function must be let go
    // before the fixture tears the scheduler down)
    struct Release
    {
        std::shared_ptr<std::atomic<bool>> flag;
        ~Release() { flag->store(true); }
    } letGo{ release };
    registerTestFunction("reap", "busy", [release](auto*, int, int, auto) {
        for (int waited = 0; !release->load() && waited < 10000; waited += 2) {
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
        }
        return 0;
    });
    auto idle = faabric::util::batchExecFactory("reap", "idle", 2);
    auto busy = faabric::util::batchExecFactory("reap", "busy", 1);
    f.plannerCli.callFunctions(idle);
    f.awaitBatch(idle);
    f.plannerCli.callFunctions(busy);
    // (dispatch to the host is asynchronous: wait for the executor to appear)
    for (int i = 0; i < 400 && f.sch.getFunctionExecutorCount(busy->messages(0)) < 1; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(idle->messages(0)), 2);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(busy->messages(0)), 1);
    // nothing is stale yet
    REQUIRE_EQ(f.sch.reapStaleExecutors(), 0);
    std::this_thread::sleep_for(std::chrono::milliseconds(2000));
    REQUIRE_EQ(f.sch.reapStaleExecutors(), 2);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(idle->messages(0)), 0);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(busy->messages(0)), 1);
    release->store(true);
    f.awaitBatch(busy);
    f.conf.reset();
}

TEST_CASE("executor case: threads forked by a function that was itself chained", "[executor][cases]")
{
    ClusterFixture f(8);
    std::atomic<int> threadsRun{ 0 };
    registerTestFunction("chain", "leaf-threads", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            threadsRun++;
            return m.appidx();
        }
        // the chained function forks three threads and adds up what they return
        auto threads = faabric::util::batchExecFactory("chain", "leaf-threads", 3);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < 3; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
        }
        threads->set_singlehosthint(true);
        int sum = 0;
        for (auto& [id, rv] : exec->executeThreads(threads, {})) {
            sum += rv;
        }
        m.set_outputdata(std::to_string(sum));
        return 0;
    });
    registerTestFunction("chain", "root", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        // (a chained call is an app of its own; only threads join their parent's)
        auto chained = faabric::util::batchExecFactory("chain", "leaf-threads", 1);
        exec->addChainedMessage(chained->messages(0));
        faabric::planner::getPlannerClient().callFunctions(chained);
        auto res = faabric::planner::getPlannerClient().getMessageResult(chained->messages(0), 5000);
        m.set_outputdata("threads said " + res.outputdata());
        return res.returnvalue();
    });
    auto req = faabric::util::batchExecFactory("chain", "root", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0), 10000);
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::string("threads said 6"));
    REQUIRE_EQ(threadsRun.load(), 3);
}
