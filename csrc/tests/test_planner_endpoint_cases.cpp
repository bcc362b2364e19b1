// One case per case (and section) of the reference's planner HTTP endpoint
// tests: exact status codes and response bodies
// (reference: tests/test/planner/test_planner_endpoint.cpp:57-860,
// src/planner/PlannerEndpointHandler.cpp:20-420)
#include "fixtures.h"

#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/planner/PlannerEndpointHandler.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/json.h>

#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <thread>

using namespace tests;
using namespace faabric::planner;

namespace {
struct EndpointFixture : ClusterFixture
{
    faabric::endpoint::FaabricEndpoint endpoint;
    int port = 0;

    explicit EndpointFixture(int slots = 8)
      : ClusterFixture(slots)
      , endpoint(0, 2, std::make_shared<PlannerEndpointHandler>())
    {
        endpoint.start(faabric::endpoint::EndpointMode::BG_THREAD);
        port = endpoint.getPort();
    }

    ~EndpointFixture()
    {
        endpoint.stop();
        planner.setPolicy("bin-pack");
    }

    std::pair<int, std::string> post(const std::string& body)
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
        std::string req = "POST / HTTP/1.1\r\nHost: localhost\r\nContent-Length: " + std::to_string(body.size()) +
                          "\r\nConnection: close\r\n\r\n" + body;
        ::send(fd, req.data(), req.size(), 0);
        std::string resp;
        char buf[8192];
        ssize_t n;
        while ((n = ::recv(fd, buf, sizeof(buf), 0)) > 0) {
            resp.append(buf, (size_t)n);
        }
        ::close(fd);
        size_t hdrEnd = resp.find("\r\n\r\n");
        return { std::atoi(resp.c_str() + 9), hdrEnd == std::string::npos ? "" : resp.substr(hdrEnd + 4) };
    }

    std::pair<int, std::string> send(HttpMessage::Type type, const std::string& payload = "")
    {
        HttpMessage m;
        m.set_type(type);
        if (!payload.empty()) {
            m.set_payloadjson(payload);
        }
        return post(faabric::util::messageToJson(m));
    }

    void registerFoo()
    {
        auto reg = std::make_shared<RegisterHostRequest>();
        reg->mutable_host()->set_ip("foo");
        reg->mutable_host()->set_slots(12);
        plannerCli.registerHost(reg);
    }

    // polls EXECUTE_BATCH_STATUS until the batch is finished
    faabric::BatchExecuteRequestStatus waitFor(const std::shared_ptr<faabric::BatchExecuteRequest>& ber)
    {
        auto asked = faabric::util::batchExecStatusFactory(ber->appid());
        asked->set_expectednummessages(ber->messages_size());
        faabric::BatchExecuteRequestStatus now;
        for (int i = 0; i < 1000; i++) {
            auto [code, body] = send(HttpMessage::EXECUTE_BATCH_STATUS, faabric::util::messageToJson(*asked));
            if (code == 200) {
                faabric::util::jsonToMessage(body, &now);
                if (now.finished()) {
                    return now;
                }
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
        throw std::runtime_error("batch did not finish");
    }
};
}

TEST_CASE("endpoint case: planner reset", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::RESET);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Planner fully reset!"));
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
    f.registerFoo();
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 1u);
    r = f.send(HttpMessage::RESET);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Planner fully reset!"));
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
}

TEST_CASE("endpoint case: flushing the available hosts", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::FLUSH_AVAILABLE_HOSTS);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Flushed available hosts!"));
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
    f.registerFoo();
    REQUIRE_EQ(f.plannerCli.getAvailableHosts().size(), 1u);
    r = f.send(HttpMessage::FLUSH_AVAILABLE_HOSTS);
    REQUIRE_EQ(r.second, std::string("Flushed available hosts!"));
    REQUIRE(f.plannerCli.getAvailableHosts().empty());
}

TEST_CASE("endpoint case: flushing executors reaches every registered host", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto req = faabric::util::batchExecFactory("foo", "bar", 2);
    f.plannerCli.callFunctions(req);
    f.awaitBatch(req);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 2);
    auto r = f.send(HttpMessage::FLUSH_EXECUTORS);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Flushed executors!"));
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 0);
    REQUIRE_EQ(f.factory->flushCount, 1);
}

TEST_CASE("endpoint case: the available hosts, before and after registrations", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    f.send(HttpMessage::RESET);
    auto r = f.send(HttpMessage::GET_AVAILABLE_HOSTS);
    REQUIRE_EQ(r.first, 200);
    AvailableHostsResponse none;
    faabric::util::jsonToMessage(r.second, &none);
    REQUIRE_EQ(none.hosts_size(), 0);
    auto reg = std::make_shared<RegisterHostRequest>();
    for (auto [ip, slots] : std::vector<std::pair<std::string, int>>{ { "foo", 12 }, { "bar", 4 } }) {
        reg->mutable_host()->set_ip(ip);
        reg->mutable_host()->set_slots(slots);
        f.plannerCli.registerHost(reg);
    }
    r = f.send(HttpMessage::GET_AVAILABLE_HOSTS);
    REQUIRE_EQ(r.first, 200);
    AvailableHostsResponse two;
    faabric::util::jsonToMessage(r.second, &two);
    REQUIRE_EQ(two.hosts_size(), 2);
    std::map<std::string, int> got;
    for (const auto& h : two.hosts()) {
        got[h.ip()] = h.slots();
    }
    REQUIRE_EQ(got["foo"], 12);
    REQUIRE_EQ(got["bar"], 4);
}

TEST_CASE("endpoint case: the planner config", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::GET_CONFIG);
    REQUIRE_EQ(r.first, 200);
    PlannerConfig cfg;
    faabric::util::jsonToMessage(r.second, &cfg);
    REQUIRE(!cfg.ip().empty());
    REQUIRE(cfg.hosttimeout() > 0);
    REQUIRE(cfg.numthreadshttpserver() > 0);
}

TEST_CASE("endpoint case: the execution graph of a finished message", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto ber = faabric::util::batchExecFactory("foo", "bar", 1);
    f.plannerCli.callFunctions(ber);
    auto result = f.plannerCli.getMessageResult(ber->appid(), ber->messages(0).id(), 2000);
    auto r = f.send(HttpMessage::GET_EXEC_GRAPH, faabric::util::messageToJson(ber->messages(0)));
    REQUIRE_EQ(r.first, 200);
    faabric::util::ExecGraph expected{ .rootNode = faabric::util::ExecGraphNode{ .msg = result } };
    REQUIRE_EQ(r.second, faabric::util::execGraphToJson(expected));
}

TEST_CASE("endpoint case: the execution graph of an unknown app fails", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto ber = faabric::util::batchExecFactory("foo", "bar", 1);
    f.plannerCli.callFunctions(ber);
    f.plannerCli.getMessageResult(ber->appid(), ber->messages(0).id(), 2000);
    faabric::Message other = ber->messages(0);
    other.set_appid(1337);
    auto r = f.send(HttpMessage::GET_EXEC_GRAPH, faabric::util::messageToJson(other));
    REQUIRE_EQ(r.first, 500);
    REQUIRE_EQ(r.second, std::string("Failed getting exec. graph!"));
}

TEST_CASE("endpoint case: the execution graph request needs a message as payload", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::GET_EXEC_GRAPH, "foo bar");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in request body"));
}

TEST_CASE("endpoint case: executing a batch answers with its status", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto ber = faabric::util::batchExecFactory("foo", "bar", 3);
    auto r = f.send(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*ber));
    REQUIRE_EQ(r.first, 200);
    faabric::BatchExecuteRequestStatus status;
    faabric::util::jsonToMessage(r.second, &status);
    REQUIRE_EQ(status.appid(), ber->appid());
    REQUIRE_EQ(status.expectednummessages(), 3);
    auto done = f.waitFor(ber);
    REQUIRE_EQ(done.messageresults_size(), 3);
}

TEST_CASE("endpoint case: a batch with a bad payload, an inconsistent batch, a batch that does not fit", "[planner][endpoint][cases]")
{
    EndpointFixture f(2);
    auto r = f.send(HttpMessage::EXECUTE_BATCH, "foo bar");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in body's payload"));
    auto invalid = faabric::util::batchExecFactory("foo", "bar", 2);
    invalid->mutable_messages(1)->set_appid(1337);
    r = f.send(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*invalid));
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad BatchExecRequest"));
    auto tooBig = faabric::util::batchExecFactory("foo", "bar", 10);
    r = f.send(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*tooBig));
    REQUIRE_EQ(r.first, 500);
    REQUIRE_EQ(r.second, std::string("No available hosts"));
}

TEST_CASE("endpoint case: the status of a batch: unknown app, bad payload, finished", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto unknown = faabric::util::batchExecStatusFactory(1337);
    auto r = f.send(HttpMessage::EXECUTE_BATCH_STATUS, faabric::util::messageToJson(*unknown));
    REQUIRE_EQ(r.first, 500);
    REQUIRE_EQ(r.second, std::string("App not registered in results"));
    r = f.send(HttpMessage::EXECUTE_BATCH_STATUS, "foo bar");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in request body"));
    auto ber = faabric::util::batchExecFactory("foo", "bar", 2);
    f.plannerCli.callFunctions(ber);
    auto done = f.waitFor(ber);
    REQUIRE(done.finished());
    REQUIRE_EQ(done.appid(), ber->appid());
}

TEST_CASE("endpoint case: flushing the scheduling state forgets in-flight apps", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto release = std::make_shared<std::atomic<bool>>(false);
    registerTestFunction("foo", "hold", [release](auto*, int, int, auto) {
        for (int waited = 0; !release->load() && waited < 10000; waited += 1) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    auto ber = faabric::util::batchExecFactory("foo", "hold", 2);
    f.plannerCli.callFunctions(ber);
    REQUIRE_EQ(f.planner.getInFlightReqs().size(), 1u);
    auto r = f.send(HttpMessage::FLUSH_SCHEDULING_STATE);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Flushed scheduling state!"));
    REQUIRE_EQ(f.planner.getInFlightReqs().size(), 0u);
    release->store(true);
    std::this_thread::sleep_for(std::chrono::milliseconds(50));
}

TEST_CASE("endpoint case: in-flight apps before, during and after a batch, and the next evicted VM", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    GetInFlightAppsResponse expected;
    expected.set_nummigrations(0);
    auto r = f.send(HttpMessage::GET_IN_FLIGHT_APPS);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, faabric::util::messageToJson(expected));
    auto release = std::make_shared<std::atomic<bool>>(false);
    registerTestFunction("foo", "bar5", [release](auto*, int, int, auto) {
        for (int waited = 0; !release->load() && waited < 10000; waited += 1) {
            std::this_thread::sleep_for(std::chrono::milliseconds(1));
        }
        return 0;
    });
    auto ber = faabric::util::batchExecFactory("foo", "bar5", 5);
    r = f.send(HttpMessage::EXECUTE_BATCH, faabric::util::messageToJson(*ber));
    REQUIRE_EQ(r.first, 200);
    auto* app = expected.add_apps();
    app->set_appid(ber->appid());
    for (int i = 0; i < 5; i++) {
        app->add_hostips(f.conf.endpointHost);
    }
    r = f.send(HttpMessage::GET_IN_FLIGHT_APPS);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, faabric::util::messageToJson(expected));
    release->store(true);
    f.waitFor(ber);
    // with the spot policy the next evicted VMs are reported too
    REQUIRE_EQ(f.send(HttpMessage::SET_POLICY, "spot").first, 200);
    SetEvictedVmIpsRequest evicted;
    evicted.add_vmips(f.conf.endpointHost);
    REQUIRE_EQ(f.send(HttpMessage::SET_NEXT_EVICTED_VM, faabric::util::messageToJson(evicted)).first, 200);
    GetInFlightAppsResponse after;
    after.add_nextevictedvmips(f.conf.endpointHost);
    r = f.send(HttpMessage::GET_IN_FLIGHT_APPS);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, faabric::util::messageToJson(after));
}

TEST_CASE("endpoint case: pre-loading a scheduling decision, then running the batch", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto ber = faabric::util::batchExecFactory("foo", "bar", 5);
    std::string execBody = faabric::util::messageToJson(*ber);
    for (int i = 0; i < ber->messages_size(); i++) {
        ber->mutable_messages(i)->set_executedhost(f.conf.endpointHost);
        ber->mutable_messages(i)->set_groupidx(i);
    }
    auto r = f.send(HttpMessage::PRELOAD_SCHEDULING_DECISION, faabric::util::messageToJson(*ber));
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Decision pre-loaded to planner"));
    r = f.send(HttpMessage::EXECUTE_BATCH, execBody);
    REQUIRE_EQ(r.first, 200);
    auto done = f.waitFor(ber);
    REQUIRE_EQ(done.messageresults_size(), 5);
}

TEST_CASE("endpoint case: pre-loading with a bad payload", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::PRELOAD_SCHEDULING_DECISION, "foo bar");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in request body"));
}

TEST_CASE("endpoint case: setting and reading back each planner policy", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    for (const char* policy : { "bin-pack", "compact", "spot" }) {
        auto r = f.send(HttpMessage::SET_POLICY, policy);
        REQUIRE_EQ(r.first, 200);
        REQUIRE_EQ(r.second, std::string("Policy set correctly"));
        r = f.send(HttpMessage::GET_POLICY);
        REQUIRE_EQ(r.first, 200);
        REQUIRE_EQ(r.second, std::string(policy));
    }
}

TEST_CASE("endpoint case: an unknown policy name is refused", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.send(HttpMessage::SET_POLICY, "foo-bar");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Unrecognised policy name: foo-bar"));
    REQUIRE_EQ(f.send(HttpMessage::GET_POLICY).second, std::string("bin-pack"));
}

TEST_CASE("endpoint case: the next evicted VM: valid, bad body, wrong policy", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    SetEvictedVmIpsRequest evicted;
    evicted.add_vmips("1.1.1.1");
    const std::string body = faabric::util::messageToJson(evicted);
    REQUIRE_EQ(f.send(HttpMessage::SET_POLICY, "spot").first, 200);
    auto r = f.send(HttpMessage::SET_NEXT_EVICTED_VM, body);
    REQUIRE_EQ(r.first, 200);
    REQUIRE_EQ(r.second, std::string("Next evicted VM set"));
    r = f.send(HttpMessage::SET_NEXT_EVICTED_VM, "1.1.1.1");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in body's payload"));
    REQUIRE_EQ(f.send(HttpMessage::SET_POLICY, "compact").first, 200);
    r = f.send(HttpMessage::SET_NEXT_EVICTED_VM, body);
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Next evicted VM must only be set in 'spot' policy"));
}

TEST_CASE("endpoint case: empty bodies, broken JSON and unknown message types", "[planner][endpoint][cases]")
{
    EndpointFixture f;
    auto r = f.post("");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Empty request"));
    r = f.post("{\"type\": ");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Bad JSON in request body"));
    r = f.post("{\"type\": 99}");
    REQUIRE_EQ(r.first, 400);
    REQUIRE_EQ(r.second, std::string("Unrecognised message type"));
}
