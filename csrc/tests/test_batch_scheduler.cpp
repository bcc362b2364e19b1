// Table-driven policy tests.  Scenario coverage follows the reference's
// tests/test/batch-scheduler/test_{binpack,compact,spot}_scheduler.cpp:
// NEW / SCALE_CHANGE / DIST_CHANGE decisions and their tie-breaks.
#include "harness.h"

#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/batch-scheduler/DecisionCache.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>

using namespace faabric::batch_scheduler;
typedef std::vector<std::string> Hosts;

namespace {
enum Kind
{
    NEW_REQ,
    SCALE,
    DIST
};

// Sentinels for expected outcomes
const Hosts NES = { "!not-enough-slots" };
const Hosts DNM = { "!do-not-migrate" };
const Hosts FRZ = { "!must-freeze" };

struct Scenario
{
    const char* name;
    Kind kind;
    Hosts ips;
    std::vector<int> slots;
    std::vector<int> used;
    int nMsgs;
    Hosts inFlightHosts; // placement of the already-running messages
    Hosts expected;
    const char* evicted = nullptr;
};

HostMap makeHostMap(const Scenario& s)
{
    HostMap hm;
    for (size_t i = 0; i < s.ips.size(); i++) {
        hm[s.ips[i]] = std::make_shared<HostState>(s.ips[i], s.slots[i], s.used[i]);
    }
    if (s.evicted != nullptr) {
        hm.at(s.evicted)->ip = MUST_EVICT_IP;
    }
    return hm;
}

SchedulingDecision place(const std::shared_ptr<faabric::BatchExecuteRequest>& ber, const Hosts& hosts)
{
    SchedulingDecision d(ber->appid(), 0);
    REQUIRE_EQ((size_t)ber->messages_size(), hosts.size());
    for (size_t i = 0; i < hosts.size(); i++) {
        d.addMessage(hosts[i], ber->messages((int)i));
    }
    return d;
}

InFlightReqs makeInFlight(const std::shared_ptr<faabric::BatchExecuteRequest>& ber, const Hosts& hosts)
{
    InFlightReqs out;
    if (hosts.empty()) {
        return out;
    }
    std::shared_ptr<faabric::BatchExecuteRequest> old;
    if ((int)hosts.size() > ber->messages_size()) {
        old = faabric::util::batchExecFactory("bat", "man", (int)hosts.size());
    } else {
        old = faabric::util::batchExecFactory("bat", "man", 0);
        for (size_t i = 0; i < hosts.size(); i++) {
            *old->add_messages() = ber->messages((int)i);
        }
    }
    faabric::util::updateBatchExecAppId(old, ber->appid());
    out[ber->appid()] = { old, std::make_shared<SchedulingDecision>(place(old, hosts)) };
    return out;
}

void runScenarios(const std::string& mode, const std::vector<Scenario>& scenarios)
{
    faabric::util::getSystemConfig().batchSchedulerMode = mode;
    resetBatchScheduler();
    auto scheduler = getBatchScheduler();
    for (const auto& s : scenarios) {
        auto ber = faabric::util::batchExecFactory("bat", "man", s.nMsgs);
        if (s.kind == DIST) {
            ber->set_type(faabric::BatchExecuteRequest::MIGRATION);
        }
        auto hostMap = makeHostMap(s);
        auto inFlight = makeInFlight(ber, s.inFlightHosts);
        auto actual = scheduler->makeSchedulingDecision(hostMap, inFlight, ber);
        REQUIRE(actual != nullptr);
        std::string ctx = mode + " / " + s.name + ": got " + actual->toString();
        if (s.expected == NES) {
            if (!(*actual == NOT_ENOUGH_SLOTS_DECISION)) {
                fbtest::fail(__FILE__, __LINE__, ctx + " (expected NOT_ENOUGH_SLOTS)");
            }
        } else if (s.expected == DNM) {
            if (!(*actual == DO_NOT_MIGRATE_DECISION)) {
                fbtest::fail(__FILE__, __LINE__, ctx + " (expected DO_NOT_MIGRATE)");
            }
        } else if (s.expected == FRZ) {
            if (!(*actual == MUST_FREEZE_DECISION)) {
                fbtest::fail(__FILE__, __LINE__, ctx + " (expected MUST_FREEZE)");
            }
        } else {
            auto expected = place(ber, s.expected);
            bool same = actual->appId == expected.appId && actual->nFunctions == expected.nFunctions &&
                        actual->hosts == expected.hosts && actual->messageIds == expected.messageIds &&
                        actual->appIdxs == expected.appIdxs && actual->groupIdxs == expected.groupIdxs;
            if (!same) {
                fbtest::fail(__FILE__, __LINE__, ctx + " expected " + expected.toString());
            }
        }
        fbtest::assertionCount()++;
    }
    faabric::util::getSystemConfig().reset();
    resetBatchScheduler();
}

// NEW and SCALE_CHANGE behave identically in all three policies
std::vector<Scenario> commonScenarios()
{
    return {
        { "new: not enough slots", NEW_REQ, { "foo", "bar" }, { 1, 1 }, { 0, 0 }, 6, {}, NES },
        { "new: fits in one host", NEW_REQ, { "foo", "bar" }, { 4, 3 }, { 0, 0 }, 3, {}, { "foo", "foo", "foo" } },
        { "new: exactly one host", NEW_REQ, { "foo", "bar" }, { 4, 3 }, { 0, 0 }, 4, {}, { "foo", "foo", "foo", "foo" } },
        { "new: spans two hosts", NEW_REQ, { "foo", "bar" }, { 4, 3 }, { 0, 0 }, 6, {}, { "foo", "foo", "foo", "foo", "bar", "bar" } },
        { "new: exactly two hosts", NEW_REQ, { "foo", "bar" }, { 4, 3 }, { 0, 0 }, 7, {}, { "foo", "foo", "foo", "foo", "bar", "bar", "bar" } },
        { "new: more free beats larger", NEW_REQ, { "foo", "bar" }, { 3, 4 }, { 0, 2 }, 4, {}, { "foo", "foo", "foo", "bar" } },
        { "new: tie on free -> larger", NEW_REQ, { "foo", "bar" }, { 4, 3 }, { 1, 0 }, 6, {}, { "foo", "foo", "foo", "bar", "bar", "bar" } },
        { "new: full tie -> alphabetical desc", NEW_REQ, { "foo", "bar" }, { 3, 3 }, { 0, 0 }, 6, {}, { "foo", "foo", "foo", "bar", "bar", "bar" } },
        { "new: many hosts", NEW_REQ, { "foo", "bar", "baz", "bip", "bup" }, { 4, 6, 2, 3, 1 }, { 0, 2, 2, 2, 0 }, 10, {},
          { "bar", "bar", "bar", "bar", "foo", "foo", "foo", "foo", "bip", "bup" } },

        { "scale: not enough slots", SCALE, { "foo", "bar" }, { 2, 1 }, { 1, 0 }, 6, { "foo" }, NES },
        { "scale: fits in one host", SCALE, { "foo", "bar" }, { 4, 3 }, { 1, 0 }, 3, { "foo" }, { "foo", "foo", "foo" } },
        { "scale: prefers known hosts", SCALE, { "foo", "bar" }, { 5, 4 }, { 0, 1 }, 3, { "bar" }, { "bar", "bar", "bar" } },
        { "scale: spans hosts", SCALE, { "foo", "bar" }, { 4, 3 }, { 0, 1 }, 4, { "bar" }, { "bar", "bar", "foo", "foo" } },
        { "scale: more running messages first", SCALE, { "foo", "bar" }, { 4, 3 }, { 1, 2 }, 1, { "bar", "bar", "foo" }, { "bar" } },
        { "scale: known hosts first", SCALE, { "foo", "bar", "baz" }, { 4, 3, 2 }, { 0, 1, 1 }, 5, { "bar", "baz" },
          { "bar", "bar", "baz", "foo", "foo" } },
        { "scale: skips full preferred hosts", SCALE, { "foo", "bar", "baz" }, { 4, 2, 2 }, { 0, 2, 1 }, 3, { "bar", "bar", "baz" },
          { "baz", "foo", "foo" } },
        { "scale: tie -> free slots", SCALE, { "foo", "bar", "baz" }, { 4, 3, 2 }, { 0, 1, 1 }, 3, { "bar", "baz" }, { "bar", "bar", "baz" } },
        { "scale: tie -> size", SCALE, { "foo", "bar", "baz" }, { 4, 3, 2 }, { 0, 2, 1 }, 3, { "bar", "baz" }, { "bar", "baz", "foo" } },
        { "scale: tie -> alphabetical", SCALE, { "foo", "bar", "baz" }, { 4, 2, 2 }, { 0, 1, 1 }, 3, { "bar", "baz" }, { "baz", "bar", "foo" } },
    };
}
}

TEST_CASE("bin-pack: new and scale-change decisions", "[batch-scheduler]")
{
    runScenarios("bin-pack", commonScenarios());
}

TEST_CASE("compact: new and scale-change decisions", "[batch-scheduler]")
{
    runScenarios("compact", commonScenarios());
}

TEST_CASE("spot: new and scale-change decisions", "[batch-scheduler]")
{
    runScenarios("spot", commonScenarios());
}

TEST_CASE("bin-pack: dist-change (migration) decisions", "[batch-scheduler]")
{
    runScenarios(
      "bin-pack",
      {
        { "no opportunity (single host)", DIST, { "foo" }, { 4 }, { 2 }, 2, { "foo", "foo" }, DNM },
        { "no opportunity (multi host)", DIST, { "foo", "bar" }, { 4, 2 }, { 4, 1 }, 5, { "foo", "foo", "foo", "foo", "bar" }, DNM },
        { "consolidate to one host", DIST, { "foo", "bar" }, { 4, 2 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "foo", "foo", "foo", "foo" } },
        { "tie -> free slots", DIST, { "foo", "bar" }, { 4, 5 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "bar", "bar", "bar", "bar" } },
        { "tie -> capacity", DIST, { "foo", "bar" }, { 4, 5 }, { 2, 3 }, 4, { "foo", "foo", "bar", "bar" }, { "bar", "bar", "bar", "bar" } },
        { "tie -> alphabetical", DIST, { "foo", "bar" }, { 4, 4 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "foo", "foo", "foo", "foo" } },
        { "prefers hosts with more messages", DIST, { "foo", "bar", "baz" }, { 3, 2, 1 }, { 2, 1, 1 }, 4, { "foo", "foo", "bar", "baz" },
          { "foo", "foo", "bar", "foo" } },
        { "prefers hosts with more slots", DIST, { "foo", "bar", "baz" }, { 3, 5, 1 }, { 2, 1, 1 }, 4, { "foo", "foo", "bar", "baz" },
          { "bar", "bar", "bar", "bar" } },
        { "consolidates to fewer hosts", DIST, { "foo", "bar", "baz" }, { 3, 5, 1 }, { 1, 1, 1 }, 3, { "foo", "bar", "baz" }, { "bar", "bar", "bar" } },
        { "minimises cross-host links", DIST, { "foo", "bar" }, { 4, 4 }, { 3, 3 }, 6, { "foo", "foo", "foo", "bar", "bar", "bar" },
          { "foo", "foo", "foo", "bar", "bar", "foo" } },
        { "minimises links (4 hosts)", DIST, { "foo", "bar", "baz", "bat" }, { 2, 2, 1, 1 }, { 1, 1, 1, 1 }, 4, { "foo", "bar", "baz", "bat" },
          { "foo", "bar", "bar", "foo" } },
        { "moves to fresh hosts if needed", DIST, { "foo", "bar", "baz" }, { 4, 4, 4 }, { 4, 4, 0 }, 4, { "foo", "foo", "bar", "bar" },
          { "baz", "baz", "baz", "baz" } },
        { "tie prefers fewer migrations", DIST, { "foo", "bar", "baz" }, { 4, 4, 4 }, { 0, 4, 2 }, 4, { "baz", "baz", "bar", "bar" },
          { "baz", "baz", "baz", "baz" } },
        { "fewest messages moved (i)", DIST, { "foo", "bar", "baz" }, { 5, 4, 2 }, { 3, 4, 2 }, 9,
          { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "baz", "baz" }, { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "foo", "foo" } },
        { "fewest messages moved (ii)", DIST, { "foo", "bar", "baz" }, { 5, 3, 2 }, { 2, 3, 2 }, 7,
          { "bar", "bar", "bar", "baz", "baz", "foo", "foo" }, { "bar", "bar", "foo", "foo", "foo", "foo", "foo" } },
        { "fewest messages moved (iii)", DIST, { "foo", "bar", "baz" }, { 3, 3, 3 }, { 2, 3, 2 }, 7,
          { "foo", "foo", "bar", "bar", "bar", "baz", "baz" }, { "foo", "foo", "bar", "bar", "bar", "baz", "foo" } },
      });
}

TEST_CASE("compact: dist-change (migration) decisions", "[batch-scheduler]")
{
    runScenarios(
      "compact",
      {
        { "no opportunity (single host)", DIST, { "foo" }, { 4 }, { 2 }, 2, { "foo", "foo" }, DNM },
        { "no opportunity (multi host)", DIST, { "foo", "bar" }, { 4, 2 }, { 4, 1 }, 5, { "foo", "foo", "foo", "foo", "bar" }, DNM },
        { "frees one host", DIST, { "foo", "bar", "baz" }, { 4, 4, 4 }, { 2, 2, 4 }, 4, { "baz", "baz", "baz", "baz" }, { "bar", "bar", "foo", "foo" } },
        { "frees many hosts", DIST, { "foo", "bar", "baz", "lol" }, { 4, 4, 2, 2 }, { 2, 2, 2, 2 }, 4, { "baz", "baz", "lol", "lol" },
          { "bar", "bar", "foo", "foo" } },
        { "frees what it can", DIST, { "foo", "bar", "baz", "lol" }, { 4, 4, 2, 2 }, { 4, 2, 2, 2 }, 4, { "baz", "baz", "lol", "lol" },
          { "bar", "bar", "lol", "lol" } },
        { "consolidate to one host", DIST, { "foo", "bar" }, { 4, 2 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "foo", "foo", "foo", "foo" } },
        { "tie -> free slots", DIST, { "foo", "bar" }, { 4, 5 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "bar", "bar", "bar", "bar" } },
        { "tie -> capacity", DIST, { "foo", "bar" }, { 4, 5 }, { 2, 3 }, 4, { "foo", "foo", "bar", "bar" }, { "bar", "bar", "bar", "bar" } },
        { "tie -> alphabetical", DIST, { "foo", "bar" }, { 4, 4 }, { 2, 2 }, 4, { "foo", "foo", "bar", "bar" }, { "foo", "foo", "foo", "foo" } },
        { "prefers hosts with more messages", DIST, { "foo", "bar", "baz" }, { 3, 2, 1 }, { 2, 1, 1 }, 4, { "foo", "foo", "bar", "baz" },
          { "foo", "foo", "bar", "foo" } },
        { "prefers hosts with more slots", DIST, { "foo", "bar", "baz" }, { 3, 5, 1 }, { 2, 1, 1 }, 4, { "foo", "foo", "bar", "baz" },
          { "bar", "bar", "bar", "bar" } },
        { "consolidates to fewer hosts", DIST, { "foo", "bar", "baz" }, { 3, 5, 1 }, { 1, 1, 1 }, 3, { "foo", "bar", "baz" }, { "bar", "bar", "bar" } },
        { "no extra free hosts -> nothing", DIST, { "foo", "bar" }, { 4, 4 }, { 3, 3 }, 6, { "foo", "foo", "foo", "bar", "bar", "bar" }, DNM },
        { "fewest messages moved (4 hosts)", DIST, { "foo", "bar", "baz", "bat" }, { 2, 2, 1, 1 }, { 1, 1, 1, 1 }, 4, { "foo", "bar", "baz", "bat" },
          { "foo", "bar", "bar", "foo" } },
        { "ignores empty hosts", DIST, { "foo", "bar", "baz" }, { 4, 4, 4 }, { 4, 4, 0 }, 4, { "foo", "foo", "bar", "bar" }, DNM },
        { "fewest messages moved (i)", DIST, { "foo", "bar", "baz" }, { 5, 4, 2 }, { 3, 4, 2 }, 9,
          { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "baz", "baz" }, { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "foo", "foo" } },
        { "fewest messages moved (ii)", DIST, { "foo", "bar", "baz" }, { 5, 3, 2 }, { 2, 3, 2 }, 7,
          { "bar", "bar", "bar", "baz", "baz", "foo", "foo" }, { "bar", "bar", "foo", "foo", "foo", "foo", "foo" } },
      });
}

TEST_CASE("spot: dist-change (eviction) decisions", "[batch-scheduler]")
{
    runScenarios(
      "spot",
      {
        { "new apps avoid the tainted VM", NEW_REQ, { "gpu0", "gpu1", "idle" }, { 2, 2, 0 }, { 0, 0, 0 }, 1, {}, { "gpu0" }, "gpu1" },
        { "new apps avoid the tainted VM (no room)", NEW_REQ, { "gpu0", "gpu1" }, { 2, 2 }, { 1, 0 }, 2, {}, NES, "gpu1" },
        { "no tainted VMs", DIST, { "foo" }, { 4 }, { 2 }, 2, { "foo", "foo" }, DNM },
        { "no tainted VMs (multi)", DIST, { "foo", "bar" }, { 4, 2 }, { 4, 1 }, 5, { "foo", "foo", "foo", "foo", "bar" }, DNM },
        { "ignores chances to free hosts", DIST, { "foo", "bar", "baz" }, { 4, 4, 4 }, { 2, 2, 4 }, 4, { "baz", "baz", "baz", "baz" }, DNM },
        { "freeze if tainted and no room (1 host)", DIST, { "foo" }, { 4 }, { 2 }, 2, { "foo", "foo" }, FRZ, "foo" },
        { "freeze if tainted and no room (multi)", DIST, { "foo", "bar", "baz", "lol" }, { 4, 4, 2, 2 }, { 2, 4, 2, 2 }, 4,
          { "foo", "foo", "bar", "bar" }, FRZ, "foo" },
        { "migrates off the tainted VM", DIST, { "foo", "bar", "baz", "lol" }, { 4, 4, 2, 2 }, { 2, 2, 2, 2 }, 4, { "baz", "baz", "lol", "lol" },
          { "foo", "foo", "lol", "lol" }, "baz" },
        { "tainted VM without our messages", DIST, { "foo", "bar", "baz", "lol" }, { 4, 4, 2, 2 }, { 2, 2, 2, 2 }, 4, { "baz", "baz", "lol", "lol" },
          DNM, "foo" },
        { "prefers hosts with more messages", DIST, { "foo", "bar", "baz" }, { 3, 2, 1 }, { 2, 1, 1 }, 4, { "foo", "foo", "bar", "baz" },
          { "foo", "foo", "foo", "baz" }, "bar" },
        { "fewest messages moved", DIST, { "foo", "bar", "baz" }, { 5, 4, 2 }, { 3, 4, 2 }, 9,
          { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "baz", "baz" }, { "foo", "foo", "foo", "bar", "bar", "bar", "bar", "foo", "foo" },
          "baz" },
      });
}

TEST_CASE("decision type classification", "[batch-scheduler]")
{
    auto ber = faabric::util::batchExecFactory("bat", "man", 2);
    InFlightReqs none;
    REQUIRE_EQ((int)BatchScheduler::getDecisionType(none, ber), (int)DecisionType::NEW);
    auto inFlight = makeInFlight(ber, { "foo", "foo" });
    REQUIRE_EQ((int)BatchScheduler::getDecisionType(inFlight, ber), (int)DecisionType::SCALE_CHANGE);
    ber->set_type(faabric::BatchExecuteRequest::MIGRATION);
    REQUIRE_EQ((int)BatchScheduler::getDecisionType(inFlight, ber), (int)DecisionType::DIST_CHANGE);
}

TEST_CASE("unknown scheduler mode is rejected", "[batch-scheduler]")
{
    faabric::util::getSystemConfig().batchSchedulerMode = "nope";
    resetBatchScheduler();
    REQUIRE_THROWS(getBatchScheduler());
    faabric::util::getSystemConfig().reset();
    resetBatchScheduler();
}

TEST_CASE("scheduling decision bookkeeping", "[batch-scheduler]")
{
    SchedulingDecision d(123, 345);
    auto ber = faabric::util::batchExecFactory("bat", "man", 3);
    d.addMessage("hostA", ber->messages(0));
    d.addMessage("hostB", ber->messages(1));
    d.addMessage("hostA", ber->messages(2));
    REQUIRE_EQ(d.nFunctions, 3);
    REQUIRE(!d.isSingleHost());
    REQUIRE_EQ(d.uniqueHosts().size(), 2u);
    REQUIRE_EQ(d.removeMessage(ber->messages(1).id()), 0);
    REQUIRE_EQ(d.nFunctions, 2);
    REQUIRE(d.isSingleHost());
    REQUIRE_THROWS(d.removeMessage(424242));

    // Point-to-point mappings round trip
    faabric::PointToPointMappings mappings;
    mappings.set_appid(7);
    mappings.set_groupid(8);
    for (int i = 0; i < 3; i++) {
        auto* m = mappings.add_mappings();
        m->set_host(i == 1 ? "hostB" : "hostA");
        m->set_messageid(100 + i);
        m->set_appidx(i);
        m->set_groupidx(i + 10);
        m->set_mpiport(8800 + i);
    }
    auto fromMappings = SchedulingDecision::fromPointToPointMappings(mappings);
    REQUIRE_EQ(fromMappings.appId, 7u);
    REQUIRE_EQ(fromMappings.groupId, 8);
    REQUIRE_EQ(fromMappings.nFunctions, 3);
    REQUIRE_EQ(fromMappings.hosts[1], std::string("hostB"));
    REQUIRE_EQ(fromMappings.groupIdxs[2], 12);
    REQUIRE_EQ(fromMappings.mpiPorts[2], 8802);
}

TEST_CASE("decision cache", "[batch-scheduler]")
{
    auto& cache = getSchedulingDecisionCache();
    cache.clear();
    auto ber = faabric::util::batchExecFactory("bat", "man", 3);
    REQUIRE(cache.getCachedDecision(ber) == nullptr);
    SchedulingDecision d(ber->appid(), 55);
    d.addMessage("a", ber->messages(0));
    d.addMessage("b", ber->messages(1));
    d.addMessage("a", ber->messages(2));
    cache.addCachedDecision(ber, d);
    auto hit = cache.getCachedDecision(ber);
    REQUIRE(hit != nullptr);
    REQUIRE_EQ(hit->getGroupId(), 55);
    REQUIRE(hit->getHosts() == (Hosts{ "a", "b", "a" }));
    // Different size => different entry
    auto ber2 = faabric::util::batchExecFactory("bat", "man", 2);
    faabric::util::updateBatchExecAppId(ber2, ber->appid());
    REQUIRE(cache.getCachedDecision(ber2) == nullptr);
    cache.clear();
}

// ---- one case per case / section of the reference's decision suite ----
// (reference: tests/test/batch-scheduler/test_scheduling_decisions.cpp)
namespace {
void buildDecisionCase(std::string hostA, std::string hostB, std::string hostC, bool expectSingleHost)
{
    const int appId = 123, groupId = 345;
    auto req = faabric::util::batchExecFactory("foo", "bar", 3);
    SchedulingDecision decision(appId, groupId);
    faabric::Message msgA = req->messages(0), msgB = req->messages(1), msgC = req->messages(2);
    decision.addMessage(hostB, msgA);
    decision.addMessage(hostA, msgB);
    decision.addMessage(hostC, msgC);
    REQUIRE_EQ(decision.appId, (decltype(decision.appId))appId);
    REQUIRE_EQ(decision.groupId, groupId);
    REQUIRE_EQ(decision.nFunctions, 3);
    REQUIRE(decision.messageIds == (decltype(decision.messageIds){ msgA.id(), msgB.id(), msgC.id() }));
    REQUIRE(decision.hosts == (Hosts{ hostB, hostA, hostC }));
    REQUIRE(decision.uniqueHosts() == (std::set<std::string>{ hostA, hostB, hostC }));
    REQUIRE(decision.appIdxs == (decltype(decision.appIdxs){ msgA.appidx(), msgB.appidx(), msgC.appidx() }));
    REQUIRE_EQ(decision.isSingleHost(), expectSingleHost);
    auto copy = decision;
    REQUIRE(copy == decision);
    copy.groupId = 1338;
    REQUIRE(copy != decision);
    decision.print();
}
}

TEST_CASE("decision case: built over three hosts", "[batch-scheduler][cases]")
{
    buildDecisionCase("hostA", "hostB", "hostC", false);
}

TEST_CASE("decision case: built on one remote host only", "[batch-scheduler][cases]")
{
    buildDecisionCase("hostA", "hostA", "hostA", true);
}

TEST_CASE("decision case: built on this host only", "[batch-scheduler][cases]")
{
    std::string here = faabric::util::getSystemConfig().endpointHost;
    buildDecisionCase(here, here, here, true);
}

TEST_CASE("decision case: from point-to-point mappings", "[batch-scheduler][cases]")
{
    faabric::PointToPointMappings mappings;
    mappings.set_appid(123);
    mappings.set_groupid(345);
    auto* a = mappings.add_mappings();
    a->set_host("foobar");
    a->set_messageid(222);
    a->set_appidx(2);
    a->set_groupidx(22);
    auto* b = mappings.add_mappings();
    b->set_host("bazbaz");
    b->set_messageid(333);
    b->set_appidx(3);
    b->set_groupidx(33);
    auto actual = SchedulingDecision::fromPointToPointMappings(mappings);
    REQUIRE_EQ(actual.appId, (decltype(actual.appId))123);
    REQUIRE_EQ(actual.nFunctions, 2);
    REQUIRE(actual.appIdxs == (decltype(actual.appIdxs){ 2, 3 }));
    REQUIRE(actual.groupIdxs == (decltype(actual.groupIdxs){ 22, 33 }));
    REQUIRE(actual.messageIds == (decltype(actual.messageIds){ 222, 333 }));
    REQUIRE(actual.hosts == (Hosts{ "foobar", "bazbaz" }));
}

TEST_CASE("decision case: removing messages one by one until it is empty", "[batch-scheduler][cases]")
{
    auto req = faabric::util::batchExecFactory("foo", "bar", 3);
    SchedulingDecision decision(req->appid(), req->groupid());
    decision.addMessage("foo", req->messages(0));
    decision.addMessage("bar", req->messages(1));
    decision.addMessage("baz", req->messages(2));
    decision.removeMessage(req->messages(1).id());
    REQUIRE_EQ(decision.nFunctions, 2);
    REQUIRE_EQ(decision.hosts.size(), 2u);
    REQUIRE_EQ(decision.messageIds.size(), 2u);
    REQUIRE_EQ(decision.appIdxs.size(), 2u);
    REQUIRE_EQ(decision.groupIdxs.size(), 2u);
    REQUIRE_EQ(decision.mpiPorts.size(), 2u);
    REQUIRE_THROWS(decision.removeMessage(req->messages(1).id()));
    decision.removeMessage(req->messages(0).id());
    decision.removeMessage(req->messages(2).id());
    REQUIRE_EQ(decision.nFunctions, 0);
    REQUIRE(decision.hosts.empty());
    REQUIRE(decision.messageIds.empty());
    REQUIRE(decision.appIdxs.empty());
    REQUIRE(decision.groupIdxs.empty());
}

TEST_CASE("decision cache case: a cached decision comes back with its hosts and group, until cleared", "[batch-scheduler][cases]")
{
    auto& cache = getSchedulingDecisionCache();
    cache.clear();
    auto req = faabric::util::batchExecFactory("foo", "bar", 5);
    Hosts hosts = { "alpha", "alpha", "beta", "gamma", "alpha" };
    SchedulingDecision decision(123, 345);
    for (size_t i = 0; i < hosts.size(); i++) {
        decision.addMessage(hosts[i], req->messages((int)i));
    }
    REQUIRE(cache.getCachedDecision(req) == nullptr);
    cache.addCachedDecision(req, decision);
    auto actual = cache.getCachedDecision(req);
    REQUIRE(actual != nullptr);
    REQUIRE(actual->getHosts() == hosts);
    REQUIRE_EQ(actual->getGroupId(), 345);
    cache.clear();
    REQUIRE(cache.getCachedDecision(req) == nullptr);
}

TEST_CASE("decision cache case: a decision with the wrong number of hosts is refused", "[batch-scheduler][cases]")
{
    auto& cache = getSchedulingDecisionCache();
    cache.clear();
    auto req = faabric::util::batchExecFactory("foo", "bar", 3);
    SchedulingDecision decision(123, 345);
    decision.addMessage("alpha", req->messages(0));
    decision.addMessage("alpha", req->messages(1));
    REQUIRE_THROWS(cache.addCachedDecision(req, decision));
    cache.clear();
}

TEST_CASE("decision cache case: two sizes of one function are cached side by side", "[batch-scheduler][cases]")
{
    auto& cache = getSchedulingDecisionCache();
    cache.clear();
    auto reqA = faabric::util::batchExecFactory("foo", "bar", 3);
    auto reqB = faabric::util::batchExecFactory("foo", "bar", 5);
    Hosts hostsA = { "alpha", "alpha", "beta" };
    Hosts hostsB = { "alpha", "alpha", "beta", "gamma", "gamma" };
    SchedulingDecision decisionA(123, 345), decisionB(456, 789);
    for (size_t i = 0; i < hostsA.size(); i++) {
        decisionA.addMessage(hostsA[i], reqA->messages((int)i));
    }
    for (size_t i = 0; i < hostsB.size(); i++) {
        decisionB.addMessage(hostsB[i], reqB->messages((int)i));
    }
    cache.addCachedDecision(reqA, decisionA);
    cache.addCachedDecision(reqB, decisionB);
    auto actualA = cache.getCachedDecision(reqA);
    auto actualB = cache.getCachedDecision(reqB);
    REQUIRE(actualA != nullptr && actualB != nullptr);
    REQUIRE(actualA->getHosts() == hostsA);
    REQUIRE_EQ(actualA->getGroupId(), 345);
    REQUIRE(actualB->getHosts() == hostsB);
    REQUIRE_EQ(actualB->getGroupId(), 789);
    cache.clear();
}
