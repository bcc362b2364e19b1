// One case per case of the reference's scheduler tests
// (reference: tests/test/scheduler/test_scheduler.cpp:129-830,
// tests/test/scheduler/test_function_client_server.cpp:44-163)
#include "fixtures.h"

#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/scheduler/FunctionCallServer.h>
#include <faabric/mpi/MpiWorld.h>
#include <faabric/util/PeriodicBackgroundThread.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/state/State.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/gids.h>
#include <faabric/util/testing.h>

#include <thread>

using namespace tests;

TEST_CASE("scheduler case: the function call server takes its thread count from the config", "[scheduler][cases]")
{
    auto& conf = faabric::util::getSystemConfig();
    int before = conf.functionServerThreads;
    conf.functionServerThreads = 6;
    {
        faabric::scheduler::FunctionCallServer server;
        REQUIRE_EQ(server.getNThreads(), 6);
    }
    conf.functionServerThreads = before;
}

TEST_CASE("scheduler case: a flush message clears executors and state and reaches the factory", "[scheduler][cases]")
{
    ClusterFixture f(4);
    REQUIRE_EQ(f.factory->flushCount, 0);
    auto& state = faabric::state::getGlobalState();
    state.getKV("demo", "blah", 10);
    state.getKV("other", "foo", 30);
    REQUIRE_EQ(state.getKVCount(), 2u);
    faabric::util::setTestMode(true);
    auto reqA = faabric::util::batchExecFactory("dummy", "foo", 1);
    auto reqB = faabric::util::batchExecFactory("dummy", "bar", 1);
    f.plannerCli.callFunctions(reqA);
    f.awaitResult(reqA->messages(0), 2000);
    f.plannerCli.callFunctions(reqB);
    f.awaitResult(reqB->messages(0), 2000);
    auto recorded = f.sch.getRecordedMessages();
    REQUIRE_EQ(recorded.size(), 2u);
    REQUIRE_EQ(recorded[0].function(), std::string("foo"));
    REQUIRE_EQ(recorded[1].function(), std::string("bar"));
    f.sch.clearRecordedMessages();
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqA->messages(0)), 1);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqB->messages(0)), 1);
    f.awaitBatch(reqA);
    f.awaitBatch(reqB);
    // (synchronous)
    faabric::scheduler::getFunctionCallClient(f.conf.endpointHost)->sendFlush();
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqA->messages(0)), 0);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqB->messages(0)), 0);
    REQUIRE_EQ(state.getKVCount(), 0u);
    REQUIRE_EQ(f.factory->flushCount, 1);
    faabric::util::setTestMode(false);
}

TEST_CASE("scheduler case: a batch sent through the function call client runs every message", "[scheduler][cases]")
{
    const int nCalls = 30;
    ClusterFixture f(nCalls);
    faabric::util::setTestMode(true);
    f.sch.clearRecordedMessages();
    auto req = faabric::util::batchExecFactory("foo", "bar", nCalls);
    for (int i = 0; i < nCalls; i++) {
        req->mutable_messages(i)->set_executedhost(f.conf.endpointHost);
    }
    // the planner is bypassed for scheduling: it must still see the slots as used
    faabric::HostResources res;
    res.set_slots(nCalls);
    res.set_usedslots(nCalls);
    f.sch.setThisHostResources(res);
    f.sch.addHostToGlobalSet();
    // (a request handed to a host in this process is not copied: from here on
    // it belongs to the executing side)
    std::vector<faabric::Message> sent(req->messages().begin(), req->messages().end());
    faabric::scheduler::getFunctionCallClient(f.conf.endpointHost)->executeFunctions(req);
    for (const auto& m : sent) {
        REQUIRE_EQ(f.awaitResult(m, 5000).returnvalue(), 0);
    }
    REQUIRE_EQ((int)f.sch.getRecordedMessages().size(), nCalls);
    faabric::util::setTestMode(false);
}

TEST_CASE("scheduler case: the planner's result notification wakes a waiting client", "[scheduler][cases]")
{
    ClusterFixture f(2);
    auto msg = faabric::util::messageFactory("foo", "bar");
    std::atomic<int> got{ -1 };
    std::thread waiter([&] { got = f.plannerCli.getMessageResult(msg, 3000).returnvalue(); });
    std::this_thread::sleep_for(std::chrono::milliseconds(200));
    auto result = std::make_shared<faabric::Message>(msg);
    result->set_returnvalue(1337);
    faabric::scheduler::getFunctionCallClient(f.conf.endpointHost)->setMessageResult(result);
    waiter.join();
    REQUIRE_EQ(got.load(), 1337);
}

TEST_CASE("scheduler case: reset clears the executors", "[scheduler][cases]")
{
    const int nCores = 5;
    ClusterFixture f(nCores);
    auto req = faabric::util::batchExecFactory("blah", "foo", nCores);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 0);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), nCores);
    f.sch.reset();
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 0);
}

TEST_CASE("scheduler case: test mode records executed messages in order, normal mode does not", "[scheduler][cases]")
{
    ClusterFixture f(4);
    auto reqA = faabric::util::batchExecFactory("demo", "echo", 1);
    auto reqB = faabric::util::batchExecFactory("demo", "echo", 1);
    auto reqC = faabric::util::batchExecFactory("demo", "echo", 1);
    faabric::util::setTestMode(false);
    f.sch.clearRecordedMessages();
    f.plannerCli.callFunctions(reqA);
    f.awaitResult(reqA->messages(0));
    REQUIRE(f.sch.getRecordedMessages().empty());
    faabric::util::setTestMode(true);
    auto reqA2 = faabric::util::batchExecFactory("demo", "echo", 1);
    for (auto& r : { reqA2, reqB, reqC }) {
        f.plannerCli.callFunctions(r);
        f.awaitResult(r->messages(0));
    }
    auto recorded = f.sch.getRecordedMessages();
    REQUIRE_EQ(recorded.size(), 3u);
    REQUIRE_EQ(recorded[0].id(), reqA2->messages(0).id());
    REQUIRE_EQ(recorded[1].id(), reqB->messages(0).id());
    REQUIRE_EQ(recorded[2].id(), reqC->messages(0).id());
    faabric::util::setTestMode(false);
}

TEST_CASE("scheduler case: a result set by hand comes back unchanged", "[scheduler][cases]")
{
    ClusterFixture f(1);
    auto req = faabric::util::batchExecFactory("some user", "my func", 1);
    auto& msg = *req->mutable_messages(0);
    msg.set_inputdata("blahblah");
    msg.set_executedhost(f.conf.endpointHost);
    faabric::HostResources res;
    res.set_slots(1);
    res.set_usedslots(1);
    f.sch.setThisHostResources(res);
    f.sch.addHostToGlobalSet();
    f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(msg));
    faabric::Message got = f.plannerCli.getMessageResult(msg, 1000);
    REQUIRE_EQ(got.id(), msg.id());
    REQUIRE_EQ(got.appid(), msg.appid());
    REQUIRE_EQ(got.user(), std::string("some user"));
    REQUIRE_EQ(got.function(), std::string("my func"));
    REQUIRE_EQ(got.inputdata(), std::string("blahblah"));
    REQUIRE_EQ(got.executedhost(), f.conf.endpointHost);
    REQUIRE_EQ(got.returnvalue(), 0);
}

TEST_CASE("scheduler case: many threads submit and await their own batches", "[scheduler][cases]")
{
    const int nWaiters = 10, perWaiter = 4;
    ClusterFixture f(nWaiters * perWaiter);
    std::atomic<int> ok{ 0 };
    std::vector<std::thread> waiters;
    for (int w = 0; w < nWaiters; w++) {
        waiters.emplace_back([&] {
            auto& cli = faabric::planner::getPlannerClient();
            auto req = faabric::util::batchExecFactory("demo", "echo", perWaiter);
            cli.callFunctions(req);
            for (const auto& m : req->messages()) {
                if (cli.getMessageResult(req->appid(), m.id(), 5000).returnvalue() == 0) {
                    ok++;
                }
            }
        });
    }
    for (auto& t : waiters) {
        t.join();
    }
    REQUIRE_EQ(ok.load(), nWaiters * perWaiter);
}

TEST_CASE("scheduler case: chained calls logged on a message show up in its result", "[scheduler][cases]")
{
    ClusterFixture f(8);
    auto ber = faabric::util::batchExecFactory("demo", "echo", 4);
    faabric::Message& msg = *ber->mutable_messages(0);
    const faabric::Message& chainedA = ber->messages(1);
    const faabric::Message& chainedB = ber->messages(2);
    const faabric::Message& chainedC = ber->messages(3);
    faabric::HostResources res;
    res.set_slots(8);
    res.set_usedslots(4);
    f.sch.setThisHostResources(res);
    f.sch.addHostToGlobalSet();
    msg.set_executedhost(f.conf.endpointHost);
    f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(msg));
    REQUIRE(faabric::util::getChainedFunctions(msg).empty());
    // (a fresh id each time: results are set once per message)
    msg.set_id(faabric::util::generateGid());
    faabric::util::logChainedFunction(msg, chainedA);
    f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(msg));
    REQUIRE(faabric::util::getChainedFunctions(msg) == (std::set<unsigned int>{ (unsigned int)chainedA.id() }));
    msg.set_id(faabric::util::generateGid());
    faabric::util::logChainedFunction(msg, chainedA);
    faabric::util::logChainedFunction(msg, chainedB);
    faabric::util::logChainedFunction(msg, chainedC);
    f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(msg));
    REQUIRE(faabric::util::getChainedFunctions(msg) ==
            (std::set<unsigned int>{ (unsigned int)chainedA.id(), (unsigned int)chainedB.id(), (unsigned int)chainedC.id() }));
}

TEST_CASE("scheduler case: a thread result set on a remote host is pushed to the main host, with and without diffs",
          "[scheduler][cases]")
{
    for (bool withDiffs : { false, true }) {
        ClusterFixture f(2);
        faabric::util::setMockMode(true);
        faabric::snapshot::clearMockSnapshotRequests();
        faabric::Message msg = faabric::util::messageFactory("foo", "bar");
        msg.set_mainhost("otherHost");
        msg.set_executedhost(f.conf.endpointHost);
        auto exec = std::make_shared<TestExecutor>(msg);
        std::string key;
        std::vector<uint8_t> data = { 1, 2, 3 };
        std::vector<faabric::util::SnapshotDiff> diffs;
        if (withDiffs) {
            key = "foobar123";
            diffs.emplace_back(faabric::util::SnapshotDataType::Raw, faabric::util::SnapshotMergeOperation::Bytewise, 123, data);
        }
        exec->setThreadResult(msg, 123, key, diffs);
        auto pushed = faabric::snapshot::getThreadResults();
        REQUIRE_EQ(pushed.size(), 1u);
        REQUIRE_EQ(pushed[0].first, std::string("otherHost"));
        REQUIRE_EQ(std::get<0>(pushed[0].second), msg.id());
        REQUIRE_EQ(std::get<1>(pushed[0].second), 123);
        REQUIRE_EQ(std::get<2>(pushed[0].second), key);
        REQUIRE_EQ(std::get<3>(pushed[0].second), (int)diffs.size());
        exec->shutdown();
        faabric::snapshot::clearMockSnapshotRequests();
        faabric::util::setMockMode(false);
    }
}

TEST_CASE("scheduler case: executors are reused by the next batch of the same function", "[scheduler][cases]")
{
    ClusterFixture f(4);
    auto reqA = faabric::util::batchExecFactory("foo", "bar", 2);
    auto reqB = faabric::util::batchExecFactory("foo", "bar", 2);
    f.plannerCli.callFunctions(reqA);
    for (const auto& m : reqA->messages()) {
        REQUIRE_EQ(f.awaitResult(m).returnvalue(), 0);
    }
    f.awaitBatch(reqA);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqA->messages(0)), 2);
    f.plannerCli.callFunctions(reqB);
    for (const auto& m : reqB->messages()) {
        REQUIRE_EQ(f.awaitResult(m).returnvalue(), 0);
    }
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqA->messages(0)), 2);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(reqB->messages(0)), 2);
}

TEST_CASE("scheduler case: point-to-point mappings of a decision are set here and sent to the other host", "[scheduler][cases]")
{
    ClusterFixture f(2);
    faabric::util::setMockMode(true);
    faabric::transport::clearSentMessages();
    const std::string thisHost = f.conf.endpointHost;
    const std::string otherHost = "10.0.0.1"; // (sorts after this host: ties go to the larger address)
    auto other = std::make_shared<faabric::HostResources>();
    other->set_slots(2);
    f.sch.addHostToGlobalSet(otherHost, other);
    auto req = faabric::util::batchExecFactory("foo", "bar", 4);
    for (int i = 0; i < 4; i++) {
        req->mutable_messages(i)->set_groupidx(i);
    }
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, 4);
    REQUIRE_EQ(decision.hosts, (std::vector<std::string>{ thisHost, thisHost, otherHost, otherHost }));
    REQUIRE_EQ(decision.groupIdxs, (std::vector<int>{ 0, 1, 2, 3 }));
    auto& broker = faabric::transport::getPointToPointBroker();
    REQUIRE_EQ(broker.getIdxsRegisteredForGroup(decision.groupId).size(), 4u);
    auto sent = faabric::transport::getSentMappings();
    REQUIRE_EQ(sent.size(), 1u);
    REQUIRE_EQ(sent[0].first, otherHost);
    faabric::util::setMockMode(false);
    f.planner.reset();
    broker.clear();
}

TEST_CASE("scheduler case: a transport message cached with a thread result outlives its sender's scope", "[scheduler][cases]")
{
    ClusterFixture f(1);
    REQUIRE_EQ(f.sch.getCachedMessageCount(), 0u);
    const uint8_t* data = nullptr;
    {
        std::vector<uint8_t> payload(100, 0);
        payload[0] = 1;
        payload[1] = 2;
        payload[2] = 3;
        faabric::transport::Message msg(0, NO_SEQUENCE_NUM, std::move(payload));
        data = msg.udata().data();
        f.sch.setThreadResultLocally(1, 123, 0, msg);
    }
    REQUIRE_EQ(f.sch.getCachedMessageCount(), 1u);
    REQUIRE(data[0] == 1 && data[1] == 2 && data[2] == 3);
}

TEST_CASE("scheduler case: a snapshot deletion is broadcast to every other host", "[scheduler][cases]")
{
    ClusterFixture f(2);
    faabric::util::setMockMode(true);
    faabric::snapshot::clearMockSnapshotRequests();
    for (const char* ip : { "10.0.0.7", "10.0.0.8" }) {
        auto res = std::make_shared<faabric::HostResources>();
        res->set_slots(1);
        f.sch.addHostToGlobalSet(ip, res);
    }
    auto msg = faabric::util::messageFactory("demo", "echo");
    msg.set_mainhost(f.conf.endpointHost);
    f.sch.broadcastSnapshotDelete(msg, "some-snapshot");
    auto deletes = faabric::snapshot::getSnapshotDeletes();
    REQUIRE_EQ(deletes.size(), 2u);
    std::set<std::string> hosts;
    for (auto& [host, key] : deletes) {
        hosts.insert(host);
        REQUIRE_EQ(key, std::string("some-snapshot"));
    }
    REQUIRE(hosts == (std::set<std::string>{ "10.0.0.7", "10.0.0.8" }));
    faabric::snapshot::clearMockSnapshotRequests();
    faabric::util::setMockMode(false);
    f.planner.reset();
}

TEST_CASE("scheduler case: legacy topology hints keep their names", "[batch-scheduler][cases]")
{
    using namespace faabric::batch_scheduler;
    REQUIRE_EQ(strToTopologyHint.size(), 5u);
    for (const auto& [name, hint] : strToTopologyHint) {
        REQUIRE_EQ(topologyHintToStr.at(hint), name);
    }
    REQUIRE(strToTopologyHint.at("NEVER_ALONE") == SchedulingTopologyHint::NEVER_ALONE);
    REQUIRE((int)MigrationStrategy::BIN_PACK != (int)MigrationStrategy::EMPTY_HOSTS);
    REQUIRE_EQ(std::string(DEFAULT_STATE_HOST), std::string(ANY_HOST));
    REQUIRE_EQ(DEFAULT_BACKGROUND_INTERVAL_SECONDS, 30);
    REQUIRE_EQ(NUM_MPI_EXEC_GRAPH_DETAILS, 2);
}

TEST_CASE("scheduler case: a join that times out reports the missing thread as failed", "[scheduler][cases]")
{
    ClusterFixture f(2);
    auto req = faabric::util::batchExecFactory("demo", "never-runs", 2);
    // one thread's result arrives, the other never does
    faabric::HostResources res;
    res.set_slots(2);
    res.set_usedslots(1);
    f.sch.setThisHostResources(res);
    auto done = std::make_shared<faabric::Message>(req->messages(0));
    done->set_returnvalue(0);
    done->set_executedhost(f.conf.endpointHost);
    f.plannerCli.setMessageResult(done);
    auto results = f.sch.awaitThreadResults(req, 300);
    REQUIRE_EQ(results.size(), 2u);
    REQUIRE_EQ(results[0].first, (uint32_t)req->messages(0).id());
    REQUIRE_EQ(results[0].second, 0);
    REQUIRE_EQ(results[1].first, (uint32_t)req->messages(1).id());
    REQUIRE(results[1].second != 0);
}
