// Wide batches: tree launch, combined result recording, per-batch executor keys
// (reference behaviour: src/scheduler/Scheduler.cpp:250-387 - one executor per
// message of a FUNCTIONS batch -, src/planner/Planner.cpp:807-1394 - every
// result recorded once, slots released, waiters notified)
#include "fixtures.h"

#include <faabric/planner/Planner.h>
#include <faabric/util/batch.h>

#include <set>
#include <thread>

using namespace tests;

TEST_CASE("fan-out case: every function of a wide batch runs exactly once, each in an executor of its own",
          "[planner][executor][fanout]")
{
    ClusterFixture f(0, 2, 100);
    std::mutex mx;
    std::map<int, int> runsOfMessage;
    std::set<faabric::executor::Executor*> executorsSeen;
    std::atomic<int> concurrent{ 0 }, peak{ 0 };
    registerTestFunction("demo", "wide", [&](auto* exec, int, int idx, auto req) {
        int now = ++concurrent;
        int seen = peak.load();
        while (now > seen && !peak.compare_exchange_weak(seen, now)) {
        }
        {
            std::lock_guard<std::mutex> lk(mx);
            runsOfMessage[req->messages(idx).id()]++;
            executorsSeen.insert(exec);
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
        --concurrent;
        return 0;
    });
    const int n = 160;
    auto req = faabric::util::batchExecFactory("demo", "wide", n);
    auto decision = f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.nFunctions, n);
    auto status = f.awaitBatch(req, 60000);
    REQUIRE_EQ(status->messageresults_size(), n);
    REQUIRE_EQ((int)runsOfMessage.size(), n);
    for (auto& [id, runs] : runsOfMessage) {
        REQUIRE_EQ(runs, 1);
    }
    // functions of one batch never share an executor while they run
    REQUIRE_EQ((int)executorsSeen.size(), n);
    // the launch really fanned out (not one after the other)
    REQUIRE(peak.load() > 8);
    // every slot came back
    for (auto& h : f.plannerCli.getAvailableHosts()) {
        REQUIRE_EQ(h.usedslots(), 0);
    }
    f.awaitBatch(req);
    // a second batch reuses the warm executors instead of creating more
    runsOfMessage.clear();
    auto again = faabric::util::batchExecFactory("demo", "wide", n);
    f.plannerCli.callFunctions(again);
    f.awaitBatch(again, 60000);
    REQUIRE_EQ((int)runsOfMessage.size(), n);
    REQUIRE_EQ((int)executorsSeen.size(), n);
}

TEST_CASE("fan-out case: a request may mix functions (chaining by name); each runs in an executor of its function",
          "[planner][executor][fanout]")
{
    // (the scheduler builds the executor key once per run of equal messages:
    // the cached key must follow when the function changes)
    ClusterFixture f(16);
    std::mutex mx;
    std::map<std::string, std::set<faabric::executor::Executor*>> executorsOf;
    auto body = [&](auto* exec, int, int idx, auto r) {
        std::lock_guard<std::mutex> lk(mx);
        executorsOf[r->messages(idx).function()].insert(exec);
        return 0;
    };
    registerTestFunction("demo", "alpha", body);
    registerTestFunction("demo", "beta", body);
    auto req = faabric::util::batchExecFactory("demo", "alpha", 8);
    for (int i = 0; i < 8; i += 2) {
        req->mutable_messages(i)->set_function("beta"); // alternate
    }
    REQUIRE(faabric::util::isBatchExecRequestValid(req));
    f.plannerCli.callFunctions(req);
    auto status = f.awaitBatch(req);
    REQUIRE_EQ(status->messageresults_size(), 8);
    REQUIRE_EQ(executorsOf["alpha"].size(), 4u);
    REQUIRE_EQ(executorsOf["beta"].size(), 4u);
    for (auto* e : executorsOf["alpha"]) {
        REQUIRE(executorsOf["beta"].count(e) == 0);
    }
    // what is NOT allowed: another user, another app, an empty function
    auto otherUser = faabric::util::batchExecFactory("demo", "alpha", 2);
    otherUser->mutable_messages(1)->set_user("somebody");
    REQUIRE(!faabric::util::isBatchExecRequestValid(otherUser));
    auto noFunction = faabric::util::batchExecFactory("demo", "alpha", 2);
    noFunction->mutable_messages(1)->set_function("");
    REQUIRE(!faabric::util::isBatchExecRequestValid(noFunction));
    REQUIRE_THROWS(f.plannerCli.callFunctions(noFunction));
}

TEST_CASE("fan-out case: results submitted concurrently are all recorded, in one piece", "[planner][fanout]")
{
    ClusterFixture f(0, 4, 64);
    // Schedule without running anything: results are fed in by hand
    faabric::util::setMockMode(true);
    const int n = 200;
    auto req = faabric::util::batchExecFactory("demo", "manual", n);
    auto decision = f.planner.callBatch(req);
    REQUIRE_EQ(decision->nFunctions, n);
    faabric::util::setMockMode(false);
    int used = 0;
    for (auto& h : f.planner.getAvailableHosts()) {
        used += h->usedslots();
    }
    REQUIRE_EQ(used, n);
    // eight producers race through the combining entry point
    std::vector<std::thread> producers;
    for (int p = 0; p < 8; p++) {
        producers.emplace_back([&, p] {
            for (int i = p; i < n; i += 8) {
                auto m = std::make_shared<faabric::Message>(req->messages(i));
                m->set_executedhost(decision->hosts.at(i));
                m->set_returnvalue(i % 7);
                m->set_outputdata("out-" + std::to_string(i));
                f.planner.submitMessageResult(m);
            }
        });
    }
    for (auto& t : producers) {
        t.join();
    }
    REQUIRE(f.planner.waitForAppToFinish(req->appid(), 5000));
    auto status = f.planner.getBatchResults(req->appid());
    REQUIRE_EQ(status->messageresults_size(), n);
    REQUIRE(status->finished());
    std::map<int, const faabric::Message*> byId;
    for (auto& m : status->messageresults()) {
        byId[m.id()] = &m;
    }
    for (int i = 0; i < n; i++) {
        auto it = byId.find(req->messages(i).id());
        REQUIRE(it != byId.end());
        REQUIRE_EQ(it->second->returnvalue(), i % 7);
        REQUIRE_EQ(it->second->outputdata(), "out-" + std::to_string(i));
    }
    // slots and in-flight bookkeeping are back to zero, duplicates do not release twice
    used = 0;
    for (auto& h : f.planner.getAvailableHosts()) {
        used += h->usedslots();
    }
    REQUIRE_EQ(used, 0);
    REQUIRE_EQ(f.planner.getInFlightReqs().size(), 0u);
    auto dup = std::make_shared<faabric::Message>(req->messages(3));
    dup->set_executedhost(decision->hosts.at(3));
    f.planner.submitMessageResult(dup);
    used = 0;
    for (auto& h : f.planner.getAvailableHosts()) {
        used += h->usedslots();
    }
    REQUIRE_EQ(used, 0);
    // the batch entry point skips migrated messages and survives an orphaned frozen one
    auto migrated = std::make_shared<faabric::Message>(req->messages(5));
    migrated->set_returnvalue(MIGRATED_FUNCTION_RETURN_VALUE);
    auto orphan = std::make_shared<faabric::Message>(faabric::util::messageFactory("demo", "nobody"));
    orphan->set_returnvalue(FROZEN_FUNCTION_RETURN_VALUE);
    f.planner.setMessageResults({ migrated, orphan });
    REQUIRE_EQ(f.planner.getBatchResults(req->appid())->messageresults_size(), n);
}
