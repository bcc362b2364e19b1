// One case per case of the reference's planner client / server tests
// (reference: tests/test/planner/test_planner_client_server.cpp:12-330)
#include "fixtures.h"

#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerClient.h>
#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/util/testing.h>

#include <thread>

using namespace tests;
using namespace faabric::planner;

namespace {
std::shared_ptr<RegisterHostRequest> registrationOf(const std::string& ip, int slots, int used = 0)
{
    auto req = std::make_shared<RegisterHostRequest>();
    req->mutable_host()->set_ip(ip);
    req->mutable_host()->set_slots(slots);
    req->mutable_host()->set_usedslots(used);
    return req;
}

// A planner with no host registered (the fixture registers this host: drop it)
struct BarePlanner : ClusterFixture
{
    BarePlanner()
      : ClusterFixture(0)
    {
        planner.reset();
        plannerCli.clearCache();
    }
};

void checkSameMessage(const faabric::Message& a, const faabric::Message& b)
{
    REQUIRE_EQ(a.id(), b.id());
    REQUIRE_EQ(a.appid(), b.appid());
    REQUIRE_EQ(a.user(), b.user());
    REQUIRE_EQ(a.function(), b.function());
    REQUIRE_EQ(a.executedhost(), b.executedhost());
    REQUIRE_EQ(a.returnvalue(), b.returnvalue());
    REQUIRE_EQ(a.outputdata(), b.outputdata());
    REQUIRE_EQ(a.finishtimestamp(), b.finishtimestamp());
}
}

TEST_CASE("planner case: ping", "[planner][cases]")
{
    BarePlanner f;
    f.plannerCli.ping(); // throws if anything is wrong
}

TEST_CASE("planner case: registering a host returns the keep-alive timeout, again and again", "[planner][cases]")
{
    BarePlanner f;
    auto reg = registrationOf("foo", 12);
    int timeout = f.plannerCli.registerHost(reg);
    REQUIRE(timeout > 0);
    REQUIRE_EQ(f.plannerCli.registerHost(reg), timeout);
}

TEST_CASE("planner case: available hosts appear on registration and expire without keep-alives", "[planner][cases]")
{
    BarePlanner f;
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
    f.planner.setHostKeepAliveTimeout(1);
    f.plannerCli.registerHost(registrationOf("foo", 12));
    auto hosts = f.plannerCli.getAvailableHosts();
    REQUIRE_EQ(hosts.size(), 1u);
    REQUIRE_EQ(hosts[0].ip(), std::string("foo"));
    REQUIRE_EQ(hosts[0].slots(), 12);
    // twice the timeout later it is gone
    std::this_thread::sleep_for(std::chrono::milliseconds(2100));
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
    f.planner.setHostKeepAliveTimeout(5);
}

TEST_CASE("planner case: removing a host", "[planner][cases]")
{
    BarePlanner f;
    f.plannerCli.registerHost(registrationOf("foo", 12));
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 1u);
    auto rem = std::make_shared<RemoveHostRequest>();
    rem->mutable_host()->set_ip("foo");
    rem->mutable_host()->set_slots(12);
    f.plannerCli.removeHost(rem);
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
    // removing it again, or a host that never was, is fine
    f.plannerCli.removeHost(rem);
}

TEST_CASE("planner case: a result is empty until it is set, then the waiting host is told", "[planner][cases]")
{
    BarePlanner f;
    faabric::util::setMockMode(true);
    faabric::scheduler::clearMockRequests();
    auto msg = std::make_shared<faabric::Message>(faabric::util::messageFactory("foo", "bar"));
    f.plannerCli.registerHost(registrationOf("foo", 12, 2));
    // not there yet: a zero timeout does not block and hands back an EMPTY message
    auto none = f.plannerCli.getMessageResult(msg->appid(), msg->id(), 0);
    REQUIRE(none.type() == faabric::Message_MessageType_EMPTY);
    // once set, the planner notifies the host that asked
    msg->set_returnvalue(1337);
    msg->set_executedhost("foo");
    f.plannerCli.setMessageResult(msg);
    std::vector<std::pair<std::string, std::shared_ptr<faabric::Message>>> told;
    for (int i = 0; i < 200 && told.empty(); i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
        told = faabric::scheduler::getMessageResults();
    }
    REQUIRE_EQ(told.size(), 1u);
    REQUIRE_EQ(told[0].first, f.conf.endpointHost);
    REQUIRE(told[0].second->type() != faabric::Message_MessageType_EMPTY);
    REQUIRE_EQ(told[0].second->id(), msg->id());
    REQUIRE_EQ(told[0].second->appid(), msg->appid());
    REQUIRE_EQ(told[0].second->returnvalue(), 1337);
    // the used slot of the host was released
    REQUIRE_EQ(f.plannerCli.getAvailableHosts()[0].usedslots(), 1);
    faabric::scheduler::clearMockRequests();
    faabric::util::setMockMode(false);
}

TEST_CASE("planner case: executing a batch of functions", "[planner][cases]")
{
    ClusterFixture f(4);
    auto req = faabric::util::batchExecFactory("foo", "bar", 4);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.messageIds.size(), 4u);
    for (int mid : decision.messageIds) {
        REQUIRE_EQ(f.plannerCli.getMessageResult(decision.appId, mid, 2000).returnvalue(), 0);
    }
}

TEST_CASE("planner case: the scheduling decision of a registered app", "[planner][cases]")
{
    ClusterFixture f(4);
    auto req = faabric::util::batchExecFactory("foo", "bar", 4);
    auto holdUntil = std::make_shared<std::atomic<bool>>(false);
    registerTestFunction("foo", "bar", [holdUntil](auto*, int, int, auto) {
        for (int waited = 0; !holdUntil->load() && waited < 10000; waited += 1) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    auto decision = f.plannerCli.callFunctions(req);
    auto same = f.plannerCli.getSchedulingDecision(req);
    REQUIRE_EQ(same.appId, decision.appId);
    REQUIRE_EQ(same.groupId, decision.groupId);
    REQUIRE_EQ(same.nFunctions, decision.nFunctions);
    REQUIRE_EQ(same.hosts, decision.hosts);
    REQUIRE_EQ(same.messageIds, decision.messageIds);
    REQUIRE_EQ(same.appIdxs, decision.appIdxs);
    REQUIRE_EQ(same.groupIdxs, decision.groupIdxs);
    holdUntil->store(true);
    for (int mid : decision.messageIds) {
        REQUIRE_EQ(f.plannerCli.getMessageResult(decision.appId, mid, 2000).returnvalue(), 0);
    }
}

TEST_CASE("planner case: the scheduling decision of an unknown app is empty", "[planner][cases]")
{
    ClusterFixture f(4);
    auto req = faabric::util::batchExecFactory("foo", "bar", 4);
    auto decision = f.plannerCli.callFunctions(req);
    for (int mid : decision.messageIds) {
        f.plannerCli.getMessageResult(decision.appId, mid, 2000);
    }
    auto other = faabric::util::batchExecFactory("foo", "bar", 4);
    faabric::util::updateBatchExecAppId(other, 1337);
    auto none = f.plannerCli.getSchedulingDecision(other);
    REQUIRE_EQ(none.appId, 0);
    REQUIRE_EQ(none.groupId, 0);
    REQUIRE_EQ(none.nFunctions, 0);
    REQUIRE(none.hosts.empty());
}

TEST_CASE("planner case: batch results are empty before the call and complete after it", "[planner][cases]")
{
    ClusterFixture f(4);
    auto req = faabric::util::batchExecFactory("foo", "bar", 4);
    auto before = f.plannerCli.getBatchResults(req);
    REQUIRE_EQ(before->appid(), 0);
    f.plannerCli.callFunctions(req);
    std::map<int, faabric::Message> results;
    for (const auto& m : req->messages()) {
        auto r = f.plannerCli.getMessageResult(req->appid(), m.id(), 2000);
        REQUIRE_EQ(r.returnvalue(), 0);
        results[r.id()] = r;
    }
    auto status = f.awaitBatch(req);
    REQUIRE_EQ(status->appid(), req->appid());
    REQUIRE_EQ(status->messageresults_size(), 4);
    REQUIRE(status->finished());
    for (const auto& m : status->messageresults()) {
        REQUIRE(results.count(m.id()) == 1);
        checkSameMessage(results[m.id()], m);
    }
}

TEST_CASE("planner case: the number of migrations starts at zero", "[planner][cases]")
{
    ClusterFixture f(4);
    REQUIRE_EQ(f.plannerCli.getNumMigrations(), 0);
}

TEST_CASE("planner case: a decision preloaded through the client is the one used", "[planner][cases]")
{
    // two hosts: bin-pack alone would fill the bigger one
    ClusterFixture f(2, 1, 8);
    auto req = faabric::util::batchExecFactory("foo", "bar", 2);
    auto preloaded = std::make_shared<faabric::batch_scheduler::SchedulingDecision>(req->appid(), req->groupid());
    for (int i = 0; i < 2; i++) {
        preloaded->addMessage(f.conf.endpointHost, 0, 0, i);
    }
    f.plannerCli.preloadSchedulingDecision(preloaded);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.hosts, (std::vector<std::string>{ f.conf.endpointHost, f.conf.endpointHost }));
    std::map<int, faabric::Message> results;
    for (const auto& m : req->messages()) {
        auto r = f.plannerCli.getMessageResult(req->appid(), m.id(), 2000);
        REQUIRE_EQ(r.returnvalue(), 0);
        REQUIRE_EQ(r.executedhost(), f.conf.endpointHost);
        results[r.id()] = r;
    }
    auto status = f.awaitBatch(req);
    REQUIRE_EQ(status->appid(), req->appid());
    for (const auto& m : status->messageresults()) {
        REQUIRE(results.count(m.id()) == 1);
        checkSameMessage(results[m.id()], m);
    }
}
