// Planner / scheduler / executor integration tests, all in one process
// (strategy: reference tests/test/planner/*.cpp, tests/test/scheduler/*.cpp,
// tests/test/executor/*.cpp, tests/test/endpoint/*.cpp)
#include "fixtures.h"

#include <numeric>

#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/planner/PlannerEndpointHandler.h>
#include <faabric/transport/common.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/gids.h>
#include <faabric/util/json.h>

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <thread>

using namespace tests;

namespace {
// Minimal blocking HTTP client for the endpoint tests
std::pair<int, std::string> httpPost(int port, const std::string& body, const std::string& method = "POST")
{
    int fd = ::socket(AF_INET, SOCK_STREAM, 0);
    sockaddr_in addr{};
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    inet_pton(AF_INET, "127.0.0.1", &addr.sin_addr);
    if (::connect(fd, (sockaddr*)&addr, sizeof(addr)) != 0) {
        ::close(fd);
        throw std::runtime_error("http connect failed");
    }
    std::string req = method + " / HTTP/1.1\r\nHost: localhost\r\nContent-Length: " + std::to_string(body.size()) +
                      "\r\nConnection: close\r\n\r\n" + body;
    ::send(fd, req.data(), req.size(), 0);
    std::string resp;
    char buf[8192];
    ssize_t n;
    while ((n = ::recv(fd, buf, sizeof(buf), 0)) > 0) {
        resp.append(buf, (size_t)n);
    }
    ::close(fd);
    int status = std::atoi(resp.c_str() + 9);
    size_t hdrEnd = resp.find("\r\n\r\n");
    return { status, hdrEnd == std::string::npos ? "" : resp.substr(hdrEnd + 4) };
}

std::string httpMsg(faabric::planner::HttpMessage::Type type, const std::string& payload = "")
{
    faabric::planner::HttpMessage m;
    m.set_type(type);
    if (!payload.empty()) {
        m.set_payloadjson(payload);
    }
    return faabric::util::messageToJson(m);
}
}

TEST_CASE("planner: host registration, keep-alive and removal", "[planner]")
{
    ClusterFixture f(4);
    auto hosts = f.plannerCli.getAvailableHosts();
    REQUIRE_EQ(hosts.size(), 1u);
    REQUIRE_EQ(hosts[0].ip(), f.conf.endpointHost);
    REQUIRE_EQ(hosts[0].slots(), 4);
    REQUIRE_EQ(hosts[0].usedslots(), 0);

    // Register two more, one twice (overwrite resets slots)
    auto req = std::make_shared<faabric::planner::RegisterHostRequest>();
    req->mutable_host()->set_ip("otherA");
    req->mutable_host()->set_slots(12);
    int timeout = f.plannerCli.registerHost(req);
    REQUIRE(timeout > 0);
    req->mutable_host()->set_slots(2);
    f.plannerCli.registerHost(req); // no overwrite: keep-alive only
    req->set_overwrite(true);
    req->mutable_host()->set_ip("otherB");
    req->mutable_host()->set_slots(3);
    f.plannerCli.registerHost(req);
    hosts = f.plannerCli.getAvailableHosts();
    REQUIRE_EQ(hosts.size(), 3u);
    std::map<std::string, int> slots;
    for (auto& h : hosts) {
        slots[h.ip()] = h.slots();
    }
    REQUIRE_EQ(slots["otherA"], 12);
    REQUIRE_EQ(slots["otherB"], 3);

    auto rm = std::make_shared<faabric::planner::RemoveHostRequest>();
    rm->mutable_host()->set_ip("otherA");
    f.plannerCli.removeHost(rm);
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 2u);

    // Hosts that stop sending keep-alives expire
    auto cfg = f.planner.getConfig();
    REQUIRE(cfg.hosttimeout() > 0);
    REQUIRE(f.planner.flush(faabric::planner::FlushType::Hosts));
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 0u);
}

TEST_CASE("planner: execute a batch and collect results", "[planner]")
{
    ClusterFixture f(8);
    registerTestFunction("demo", "square", [](auto*, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        int v = std::stoi(m.inputdata());
        m.set_outputdata(std::to_string(v * v));
        return 0;
    });
    registerTestFunction("demo", "fail", [](auto*, int, int idx, auto req) {
        req->mutable_messages(idx)->set_outputdata("went wrong");
        return 17;
    });
    registerTestFunction("demo", "throws", [](auto*, int, int, auto) -> int { throw std::runtime_error("boom"); });

    auto req = faabric::util::batchExecFactory("demo", "square", 6);
    for (int i = 0; i < 6; i++) {
        req->mutable_messages(i)->set_inputdata(std::to_string(i + 1));
    }
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, 6);
    REQUIRE(decision.isSingleHost());
    for (int i = 0; i < 6; i++) {
        auto res = f.awaitResult(req->messages(i));
        REQUIRE_EQ(res.returnvalue(), 0);
        REQUIRE_EQ(res.outputdata(), std::to_string((i + 1) * (i + 1)));
        REQUIRE_EQ(res.executedhost(), f.conf.endpointHost);
        REQUIRE(res.finishtimestamp() >= res.starttimestamp());
    }
    auto status = f.awaitBatch(req);
    REQUIRE_EQ(status->messageresults_size(), 6);
    // Slots are released again
    REQUIRE_EQ(f.plannerCli.getAvailableHosts()[0].usedslots(), 0);
    REQUIRE_EQ(f.planner.getInFlightReqs().size(), 0u);

    // Failing and throwing functions still produce results
    auto bad = faabric::util::batchExecFactory("demo", "fail", 1);
    f.plannerCli.callFunctions(bad);
    auto badRes = f.awaitResult(bad->messages(0));
    REQUIRE_EQ(badRes.returnvalue(), 17);
    REQUIRE_EQ(badRes.outputdata(), std::string("went wrong"));
    auto thrower = faabric::util::batchExecFactory("demo", "throws", 1);
    f.plannerCli.callFunctions(thrower);
    auto thrownRes = f.awaitResult(thrower->messages(0));
    REQUIRE_EQ(thrownRes.returnvalue(), 1);
    REQUIRE(thrownRes.outputdata().find("boom") != std::string::npos);

    // Too big for the cluster
    auto tooBig = faabric::util::batchExecFactory("demo", "square", 20);
    auto none = f.plannerCli.callFunctions(tooBig);
    REQUIRE_EQ((int)none.appId, NOT_ENOUGH_SLOTS);
}

TEST_CASE("planner: executors are reused and reaped", "[planner][executor]")
{
    // (set through the environment before any executor thread exists: the
    // config object is read concurrently once they run)
    setenv("BOUND_TIMEOUT", "5", 1);
    ClusterFixture f(4);
    unsetenv("BOUND_TIMEOUT");
    auto req = faabric::util::batchExecFactory("demo", "echo", 3);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 3);
    // Same function again: no new executors
    auto req2 = faabric::util::batchExecFactory("demo", "echo", 2);
    f.plannerCli.callFunctions(req2);
    f.awaitBatch(req2);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 3);
    // Reaping after the bound timeout
    REQUIRE_EQ(f.conf.boundTimeout, 5);
    std::this_thread::sleep_for(std::chrono::milliseconds(40));
    REQUIRE_EQ(f.sch.reapStaleExecutors(), 3);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 0);
}

// Executor hooks and bookkeeping (strategy: reference
// tests/test/executor/test_executor.cpp)
TEST_CASE("executor: reset after functions, not threads; no restore on one host", "[executor]")
{
    ClusterFixture f(6);
    TestExecutor::resetCount = 0;
    TestExecutor::restoreCount = 0;
    auto req = faabric::util::batchExecFactory("demo", "echo", 3);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    // the hook runs after the result is published: give it a moment
    for (int i = 0; i < 100 && TestExecutor::resetCount.load() < 3; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    REQUIRE_EQ(TestExecutor::resetCount.load(), 3);
    REQUIRE_EQ(TestExecutor::restoreCount.load(), 0);

    // A THREADS batch on the main host shares memory: neither hook runs
    std::atomic<int> ran{ 0 };
    registerTestFunction("demo", "forker", [&](auto* exec, int, int idx, auto req) {
        auto& msg = req->messages(idx);
        auto threads = faabric::util::batchExecFactory("demo", "worker-thread", 2);
        threads->set_type(faabric::BatchExecuteRequest::THREADS);
        faabric::util::updateBatchExecAppId(threads, msg.appid());
        for (int i = 0; i < 2; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
        }
        TestExecutor::resetCount = 0;
        auto results = exec->executeThreads(threads, {});
        int rc = 0;
        for (auto& [id, ret] : results) {
            rc += ret;
        }
        return rc;
    });
    registerTestFunction("demo", "worker-thread", [&](auto*, int, int, auto) {
        ran++;
        return 0;
    });
    auto forkReq = faabric::util::batchExecFactory("demo", "forker", 1);
    f.plannerCli.callFunctions(forkReq);
    auto res = f.awaitResult(forkReq->messages(0), 20000);
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(ran.load(), 2);
    REQUIRE_EQ(TestExecutor::restoreCount.load(), 0);
    // only the forker itself is reset, once it returns
    for (int i = 0; i < 100 && TestExecutor::resetCount.load() < 1; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
    REQUIRE_EQ(TestExecutor::resetCount.load(), 1);
}

TEST_CASE("executor: claims, chained messages and pool limits", "[executor]")
{
    ClusterFixture f(4);
    auto msg = faabric::util::messageFactory("demo", "direct");
    auto exec = std::make_shared<TestExecutor>(msg);
    // claim protocol
    REQUIRE(exec->tryClaim());
    REQUIRE(!exec->tryClaim());
    exec->releaseClaim();
    // (an executor counts as busy from claim to release, not only while its
    // pool threads run tasks as in the reference)
    REQUIRE(!exec->isExecuting());
    REQUIRE(exec->tryClaim());
    REQUIRE(exec->isExecuting());
    REQUIRE(exec->getMillisSinceLastExec() >= 0);

    // chained-message registry
    auto chained = faabric::util::messageFactory("demo", "child");
    chained.set_inputdata("payload");
    exec->addChainedMessage(chained);
    REQUIRE_EQ(exec->getChainedMessage(chained.id()).inputdata(), std::string("payload"));
    REQUIRE(exec->getChainedMessageIds() == (std::set<unsigned int>{ (unsigned int)chained.id() }));
    REQUIRE_THROWS(exec->getChainedMessage(chained.id() + 1));

    // memory view and growth
    size_t before = exec->getMemoryView().size();
    exec->setMemorySize(before + faabric::util::HOST_PAGE_SIZE);
    REQUIRE_EQ(exec->getMemoryView().size(), before + faabric::util::HOST_PAGE_SIZE);
    REQUIRE(exec->getMaxMemorySize() >= exec->getMemoryView().size());

    // More concurrent functions than pool threads cannot be placed (only
    // threads may double up, on pool threads that run threads)
    int pool = faabric::util::getUsableCores();
    std::atomic<bool> release{ false };
    std::atomic<int> held{ 0 };
    registerTestFunction("demo", "hold", [&](auto*, int, int, auto) {
        held++;
        while (!release.load()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    auto tooMany = faabric::util::batchExecFactory("demo", "hold", pool + 1);
    std::vector<int> idxs(pool + 1);
    std::iota(idxs.begin(), idxs.end(), 0);
    REQUIRE_THROWS(exec->executeTasks(idxs, tooMany));
    release = true;
    exec->joinThreadPool();
    REQUIRE_EQ(held.load(), pool);
    exec->shutdown();
    REQUIRE(exec->isShutdown());
}

// The remaining client verbs (strategy: reference
// tests/test/planner/test_planner_client_server.cpp)
TEST_CASE("planner client: ping, decisions in flight, migrations counter, state mains", "[planner]")
{
    ClusterFixture f(4);
    REQUIRE_NOTHROW(f.plannerCli.ping());
    REQUIRE_EQ(f.plannerCli.getNumMigrations(), 0);

    // The decision of an app can be asked for while it is in flight
    std::atomic<bool> release{ false };
    registerTestFunction("demo", "wait", [&](auto*, int, int, auto) {
        while (!release.load()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "wait", 3);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, 3);
    auto asked = f.plannerCli.getSchedulingDecision(req);
    REQUIRE_EQ(asked.appId, decision.appId);
    REQUIRE_EQ(asked.groupId, decision.groupId);
    REQUIRE(asked.hosts == decision.hosts);
    REQUIRE(asked.messageIds == decision.messageIds);
    // nothing finished yet
    auto partial = f.plannerCli.getBatchResults(req);
    REQUIRE(partial != nullptr);
    REQUIRE(!partial->finished());
    REQUIRE_EQ(partial->messageresults_size(), 0);
    release = true;
    auto status = f.awaitBatch(req);
    REQUIRE(status->finished());
    // ...and once it has left the in-flight set there is no decision any more
    auto gone = f.plannerCli.getSchedulingDecision(req);
    REQUIRE_EQ(gone.nFunctions, 0);
    // an app the planner never saw
    auto unknown = faabric::util::batchExecFactory("demo", "wait", 1);
    REQUIRE_EQ(f.plannerCli.getSchedulingDecision(unknown).nFunctions, 0);

    // State mains: first claim wins, later claimants learn the owner, dropping
    // the main lets somebody else take over
    REQUIRE_EQ(f.plannerCli.stateMain("demo", "k", "hostA", true), std::string("hostA"));
    REQUIRE_EQ(f.plannerCli.stateMain("demo", "k", "hostB", true), std::string("hostA"));
    REQUIRE_EQ(f.plannerCli.stateMain("demo", "k", "hostB", false), std::string("hostA"));
    REQUIRE_EQ(f.plannerCli.stateMain("demo", "other", "hostB", false), std::string(""));
    f.plannerCli.stateMain("demo", "k", "hostA", false, true);
    REQUIRE_EQ(f.plannerCli.stateMain("demo", "k", "hostB", true), std::string("hostB"));
}

TEST_CASE("planner: chained calls build an exec graph", "[planner]")
{
    ClusterFixture f(8);
    registerTestFunction("demo", "parent", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto chained = faabric::util::batchExecFactory("demo", "child", 2);
        faabric::util::updateBatchExecAppId(chained, m.appid());
        for (int i = 0; i < 2; i++) {
            chained->mutable_messages(i)->set_inputdata("c" + std::to_string(i));
            chained->mutable_messages(i)->set_recordexecgraph(true);
            faabric::util::logChainedFunction(m, chained->messages(i));
        }
        faabric::planner::getPlannerClient().callFunctions(chained);
        for (int i = 0; i < 2; i++) {
            auto r = faabric::planner::getPlannerClient().getMessageResult(chained->messages(i), 5000);
            if (r.returnvalue() != 0) {
                return 1;
            }
        }
        faabric::util::addDetail(m, "phase", "done");
        faabric::util::incrementCounter(m, "children", 2);
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "parent", 1);
    req->mutable_messages(0)->set_recordexecgraph(true);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0));
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.chainedmsgids_size(), 2);
    REQUIRE_EQ(res.execgraphdetails().at("phase"), std::string("done"));
    REQUIRE_EQ(res.intexecgraphdetails().at("children"), 2);

    auto graph = faabric::util::getFunctionExecGraph(res);
    REQUIRE_EQ(faabric::util::countExecGraphNodes(graph), 3);
    REQUIRE_EQ(graph.rootNode.children.size(), 2u);
    REQUIRE_EQ(faabric::util::getExecGraphHosts(graph).size(), 1u);
    std::string js = faabric::util::execGraphToJson(graph);
    REQUIRE(js.find("\"root\"") != std::string::npos);
    REQUIRE(js.find("\"chained\"") != std::string::npos);
    auto parsed = faabric::proto::JsonValue::parse(js);
    REQUIRE(parsed.find("root") != nullptr);
}

TEST_CASE("planner: preloaded decisions and policies", "[planner]")
{
    ClusterFixture f(2, 2, 4);
    auto req = faabric::util::batchExecFactory("demo", "echo", 3);
    // Preloaded placements are matched to messages by group idx
    for (int i = 0; i < 3; i++) {
        req->mutable_messages(i)->set_groupidx(i);
    }
    auto preload = std::make_shared<faabric::batch_scheduler::SchedulingDecision>(req->appid(), 0);
    preload->addMessage("gpu1", req->messages(0));
    preload->addMessage("gpu0", req->messages(1));
    preload->addMessage("gpu1", req->messages(2));
    f.plannerCli.preloadSchedulingDecision(preload);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE(decision.hosts == (std::vector<std::string>{ "gpu1", "gpu0", "gpu1" }));
    auto status = f.awaitBatch(req);
    std::map<int, std::string> executedOn;
    for (auto& m : status->messageresults()) {
        executedOn[m.id()] = m.executedhost();
    }
    REQUIRE_EQ(executedOn[req->messages(0).id()], std::string("gpu1"));
    REQUIRE_EQ(executedOn[req->messages(1).id()], std::string("gpu0"));

    REQUIRE_EQ(f.planner.getPolicy(), std::string("bin-pack"));
    f.planner.setPolicy("compact");
    REQUIRE_EQ(f.planner.getPolicy(), std::string("compact"));
    REQUIRE_THROWS(f.planner.setPolicy("nonsense"));
    REQUIRE_THROWS(f.planner.setNextEvictedVm({ "gpu0" }));
    f.planner.setPolicy("spot");
    REQUIRE_NOTHROW(f.planner.setNextEvictedVm({ "gpu0" }));
    REQUIRE(f.planner.getNextEvictedHostIps() == (std::set<std::string>{ "gpu0" }));
    f.planner.setPolicy("bin-pack");
}

TEST_CASE("planner: fan-out across virtual GPU hosts", "[planner][bench]")
{
    // The BASELINE configuration: 8 hosts, 1024 functions in one batch
    ClusterFixture f(0, 8, 128);
    registerTestFunction("bench", "noop", [](auto*, int, int, auto) { return 0; });
    auto req = faabric::util::batchExecFactory("bench", "noop", 1024);
    auto t0 = std::chrono::steady_clock::now();
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, 1024);
    REQUIRE_EQ(decision.uniqueHosts().size(), 8u);
    auto status = f.awaitBatch(req, 60000);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    REQUIRE_EQ(status->messageresults_size(), 1024);
    std::map<std::string, int> perHost;
    for (auto& m : status->messageresults()) {
        REQUIRE_EQ(m.returnvalue(), 0);
        perHost[m.executedhost()]++;
    }
    REQUIRE_EQ(perHost.size(), 8u);
    for (auto& [h, n] : perHost) {
        REQUIRE_EQ(n, 128);
    }
    printf("         1024-function fan-out over 8 hosts: %.1f ms (%.0f functions/s)\n", ms, 1024.0 / ms * 1000.0);
}

TEST_CASE("planner: HTTP endpoint", "[planner][endpoint]")
{
    ClusterFixture f(4);
    faabric::endpoint::FaabricEndpoint endpoint(0, 2, std::make_shared<faabric::planner::PlannerEndpointHandler>());
    endpoint.start(faabric::endpoint::EndpointMode::BG_THREAD);
    int port = endpoint.getPort();
    REQUIRE(port > 0);
    using faabric::planner::HttpMessage;

    auto [s0, b0] = httpPost(port, "");
    REQUIRE_EQ(s0, 400);
    REQUIRE_EQ(b0, std::string("Empty request"));
    auto [s1, b1] = httpPost(port, "{bad json");
    REQUIRE_EQ(s1, 400);
    REQUIRE_EQ(b1, std::string("Bad JSON in request body"));

    auto [s2, b2] = httpPost(port, httpMsg(HttpMessage::GET_AVAILABLE_HOSTS));
    REQUIRE_EQ(s2, 200);
    faabric::planner::AvailableHostsResponse hostsResp;
    faabric::util::jsonToMessage(b2, &hostsResp);
    REQUIRE_EQ(hostsResp.hosts_size(), 1);
    REQUIRE_EQ(hostsResp.hosts(0).slots(), 4);

    auto [s3, b3] = httpPost(port, httpMsg(HttpMessage::GET_CONFIG));
    REQUIRE_EQ(s3, 200);
    faabric::planner::PlannerConfig cfg;
    faabric::util::jsonToMessage(b3, &cfg);
    REQUIRE(cfg.hosttimeout() > 0);

    // Execute a batch and poll for its status
    auto ber = faabric::util::batchExecFactory("demo", "echo", 2);
    ber->mutable_messages(0)->set_inputdata("over http");
    auto [s4, b4] = httpPost(port, httpMsg(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*ber)));
    REQUIRE_EQ(s4, 200);
    faabric::BatchExecuteRequestStatus status;
    faabric::util::jsonToMessage(b4, &status);
    REQUIRE_EQ(status.appid(), ber->appid());
    bool finished = false;
    for (int i = 0; i < 500 && !finished; i++) {
        auto [s5, b5] = httpPost(port, httpMsg(HttpMessage::EXECUTE_BATCH_STATUS, faabric::util::messageToJson(status)));
        if (s5 == 500) {
            // Nothing has finished yet
            REQUIRE_EQ(b5, std::string("App not registered in results"));
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
            continue;
        }
        REQUIRE_EQ(s5, 200);
        faabric::BatchExecuteRequestStatus now;
        faabric::util::jsonToMessage(b5, &now);
        finished = now.finished();
        if (finished) {
            REQUIRE_EQ(now.messageresults_size(), 2);
            bool found = false;
            for (auto& m : now.messageresults()) {
                found = found || m.outputdata() == "over http";
            }
            REQUIRE(found);
        } else {
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
    }
    REQUIRE(finished);

    // Invalid batches are rejected; too-big ones report no hosts
    auto invalid = faabric::util::batchExecFactory("demo", "echo", 2);
    invalid->mutable_messages(1)->set_appid(5);
    auto [s6, b6] = httpPost(port, httpMsg(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*invalid)));
    REQUIRE_EQ(s6, 400);
    REQUIRE_EQ(b6, std::string("Bad BatchExecRequest"));
    auto tooBig = faabric::util::batchExecFactory("demo", "echo", 50);
    auto [s7, b7] = httpPost(port, httpMsg(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*tooBig)));
    REQUIRE_EQ(s7, 500);
    REQUIRE_EQ(b7, std::string("No available hosts"));
    faabric::BatchExecuteRequestStatus unknown;
    unknown.set_appid(424242);
    auto [s8, b8] = httpPost(port, httpMsg(HttpMessage::EXECUTE_BATCH_STATUS, faabric::util::messageToJson(unknown)));
    REQUIRE_EQ(s8, 500);

    // Policies, in-flight apps, flushes, reset
    auto [s9, b9] = httpPost(port, httpMsg(HttpMessage::GET_POLICY));
    REQUIRE_EQ(b9, std::string("bin-pack"));
    auto [s10, b10] = httpPost(port, httpMsg(HttpMessage::SET_POLICY, "compact"));
    REQUIRE_EQ(s10, 200);
    REQUIRE_EQ(f.planner.getPolicy(), std::string("compact"));
    auto [s11, b11] = httpPost(port, httpMsg(HttpMessage::SET_POLICY, "bogus"));
    REQUIRE_EQ(s11, 400);
    httpPost(port, httpMsg(HttpMessage::SET_POLICY, "bin-pack"));
    auto [s12, b12] = httpPost(port, httpMsg(HttpMessage::GET_IN_FLIGHT_APPS));
    REQUIRE_EQ(s12, 200);
    faabric::planner::GetInFlightAppsResponse inFlight;
    faabric::util::jsonToMessage(b12, &inFlight);
    REQUIRE_EQ(inFlight.apps_size(), 0);
    auto [s13, b13] = httpPost(port, httpMsg(HttpMessage::FLUSH_EXECUTORS));
    REQUIRE_EQ(s13, 200);
    REQUIRE_EQ(f.factory->flushCount, 1);
    auto [s14, b14] = httpPost(port, httpMsg(HttpMessage::FLUSH_SCHEDULING_STATE));
    REQUIRE_EQ(s14, 200);
    auto [s15, b15] = httpPost(port, httpMsg(HttpMessage::RESET));
    REQUIRE_EQ(s15, 200);
    REQUIRE_EQ(b15, std::string("Planner fully reset!"));
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 0u);
    auto [s16, b16] = httpPost(port, "", "OPTIONS");
    REQUIRE_EQ(s16, 200);
    endpoint.stop();
}

TEST_CASE("executor: threads share the main function's memory", "[executor][threads]")
{
    ClusterFixture f(8);
    const int nThreads = 4;
    // Each thread writes its own slot and adds into a shared Sum region
    registerTestFunction("demo", "threaded", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto mem = exec->getMemoryView();
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            int t = m.appidx();
            mem[1024 + t] = (uint8_t)(10 + t);
            // Same process: plain atomic add on the shared word
            __atomic_fetch_add((int*)(mem.data() + 64), t + 1, __ATOMIC_RELAXED);
            return t;
        }
        // Main thread
        *(int*)(mem.data() + 64) = 100;
        auto threads = faabric::util::batchExecFactory("demo", "threaded", nThreads);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < nThreads; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
        }
        threads->set_singlehost(true);
        std::vector<faabric::util::SnapshotMergeRegion> regions = { { 64,
                                                                    sizeof(int),
                                                                    faabric::util::SnapshotDataType::Int,
                                                                    faabric::util::SnapshotMergeOperation::Sum } };
        auto results = exec->executeThreads(threads, regions);
        if ((int)results.size() != nThreads) {
            return 1;
        }
        for (auto& [id, rv] : results) {
            if (rv < 1 || rv > nThreads) {
                return 2;
            }
        }
        // 100 + (2+3+4+5): app idxs are 1..4 and each adds idx+1
        int sum = *(int*)(mem.data() + 64);
        m.set_outputdata(std::to_string(sum) + ":" + std::to_string(mem[1024 + 1]) + std::to_string(mem[1024 + 4]));
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "threaded", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0));
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::string("114:1114"));
    // Everything ran in ONE executor
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 1);
    f.awaitBatch(req);
}

TEST_CASE("planner: hosts without keep-alives expire, served hosts stay", "[planner]")
{
    ClusterFixture f(2);
    int oldTimeout = f.planner.getConfig().hosttimeout();
    f.planner.setHostKeepAliveTimeout(1);
    // Outside test mode the scheduler keeps its hosts alive, including the
    // per-GPU virtual hosts it serves
    faabric::util::setTestMode(false);
    faabric::transport::registerHostAlias("gpu0", f.conf.endpointHost);
    auto res = std::make_shared<faabric::HostResources>();
    res->set_slots(3);
    f.sch.addHostToGlobalSet("gpu0", res);
    f.sch.addHostToGlobalSet();
    // A host nobody keeps alive
    auto stale = std::make_shared<faabric::planner::RegisterHostRequest>();
    stale->mutable_host()->set_ip("stale-host");
    stale->mutable_host()->set_slots(4);
    f.plannerCli.registerHost(stale);
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 3u);

    std::this_thread::sleep_for(std::chrono::milliseconds(1800));
    std::set<std::string> alive;
    for (auto& h : f.plannerCli.getAvailableHosts()) {
        alive.insert(h.ip());
    }
    REQUIRE(alive == (std::set<std::string>{ "gpu0", f.conf.endpointHost }));
    // Keep-alives did not disturb the slot accounting
    for (auto& h : f.plannerCli.getAvailableHosts()) {
        REQUIRE_EQ(h.usedslots(), 0);
        if (h.ip() == "gpu0") {
            REQUIRE_EQ(h.slots(), 3);
        }
    }
    faabric::util::setTestMode(true);
    f.planner.setHostKeepAliveTimeout(oldTimeout);
}

TEST_CASE("planner: elastic OpenMP scale-up fills the idle slots of the main host", "[planner]")
{
    ClusterFixture f(8);
    std::atomic<bool> release{ false };
    std::atomic<int> threadsRun{ 0 };
    registerTestFunction("omp", "region", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            threadsRun++;
            return 0;
        }
        // Main thread: fork 2 threads but allow the planner to scale the
        // parallel region up to whatever the host has free
        auto threads = faabric::util::batchExecFactory("omp", "region", 2);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < 2; i++) {
            auto* t = threads->mutable_messages(i);
            t->set_appidx(i + 1);
            t->set_groupidx(i + 1);
            t->set_isomp(true);
            t->set_ompnumthreads(3);
        }
        threads->set_singlehost(true);
        threads->set_singlehosthint(true);
        threads->set_elasticscalehint(true);
        auto results = exec->executeThreads(threads, {});
        m.set_outputdata(std::to_string(results.size()));
        while (!release.load()) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    // (A main message that announces its thread count gets exactly that
    // many slots preloaded; elastic growth applies to regions that were not
    // announced up-front)
    auto req = faabric::util::batchExecFactory("omp", "region", 1);
    f.plannerCli.callFunctions(req);
    // 8 slots, 1 taken by the main thread: the 2 requested threads grow to 7
    for (int i = 0; i < 2000 && threadsRun.load() < 7; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
    REQUIRE_EQ(threadsRun.load(), 7);
    release = true;
    auto res = f.awaitResult(req->messages(0));
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::string("7"));
    f.awaitBatch(req);
    REQUIRE_EQ(f.plannerCli.getAvailableHosts()[0].usedslots(), 0);
}
