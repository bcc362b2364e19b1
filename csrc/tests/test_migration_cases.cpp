// Detecting migration opportunities from inside running functions, one case
// per section of the reference's suite
// (reference: tests/test/scheduler/test_function_migration.cpp:69-260)
#include "fixtures.h"

#include <faabric/mpi/MpiWorld.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/transport/PointToPointBroker.h>

#include <thread>

using namespace tests;

namespace {
const std::string OTHER = "10.0.0.1"; // sorts below this host: ties place work here first

// This host (really executing) plus one mocked host; functions block until
// the case lets them go, so their batch stays in flight
struct MigrationCase
{
    ClusterFixture f;
    std::string mainHost;
    std::shared_ptr<std::atomic<bool>> release = std::make_shared<std::atomic<bool>>(false);

    MigrationCase(int slotsHere, int slotsOther)
      : f(slotsHere)
      , mainHost(f.conf.endpointHost)
    {
        faabric::util::setMockMode(true);
        auto other = std::make_shared<faabric::HostResources>();
        other->set_slots(slotsOther);
        f.sch.addHostToGlobalSet(OTHER, other);
        auto flag = release;
        for (const char* user : { "foo", "mpi" }) {
            registerTestFunction(user, "sleep", [flag](auto*, int, int, auto) {
                for (int i = 0; i < 2000 && !flag->load(); i++) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(5));
                }
                return 0;
            });
        }
    }

    ~MigrationCase()
    {
        release->store(true);
        faabric::util::setMockMode(false);
        faabric::transport::clearSentMessages();
        faabric::scheduler::clearMockRequests();
        f.planner.reset();
        faabric::transport::getPointToPointBroker().clear();
    }

    void updateLocalResources(int slots, int used)
    {
        faabric::HostResources r;
        r.set_slots(slots);
        r.set_usedslots(used);
        f.sch.setThisHostResources(r);
    }

    // Results of the mocked host's messages never arrive by themselves
    void finish(std::shared_ptr<faabric::BatchExecuteRequest> req, const std::vector<std::string>& executedOn)
    {
        release->store(true);
        for (int i = 0; i < req->messages_size(); i++) {
            if (executedOn[i] != mainHost) {
                req->mutable_messages(i)->set_executedhost(executedOn[i]);
                f.plannerCli.setMessageResult(std::make_shared<faabric::Message>(req->messages(i)));
            }
        }
    }
};

void twoFunctions(bool mustMigrate)
{
    MigrationCase c(1, 1);
    auto req = faabric::util::batchExecFactory("foo", "sleep", 2);
    for (int i = 0; i < 2; i++) {
        req->mutable_messages(i)->set_groupidx(i);
    }
    auto decision = c.f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.hosts, (std::vector<std::string>{ c.mainHost, OTHER }));
    for (int i = 0; i < 2; i++) {
        req->mutable_messages(i)->set_groupid(decision.groupId);
    }
    if (mustMigrate) {
        // a slot frees up next to idx 0
        c.updateLocalResources(2, 1);
    }
    auto migration0 = c.f.sch.checkForMigrationOpportunities(*req->mutable_messages(0));
    if (mustMigrate) {
        REQUIRE(migration0 != nullptr);
        REQUIRE_EQ(migration0->appid(), (int)decision.appId);
        REQUIRE(migration0->groupid() != decision.groupId);
        REQUIRE_EQ(migration0->groupidx(), 0);
        REQUIRE_EQ(migration0->srchost(), migration0->dsthost());
        REQUIRE_EQ(migration0->dsthost(), decision.hosts.at(0));
        // idx 1 learns the new group id from idx 0 (handed over by hand here)
        auto migration1 = c.f.sch.checkForMigrationOpportunities(*req->mutable_messages(1), migration0->groupid());
        REQUIRE(migration1 != nullptr);
        REQUIRE_EQ(migration1->appid(), (int)decision.appId);
        REQUIRE(migration1->groupid() != decision.groupId);
        REQUIRE_EQ(migration1->groupidx(), 1);
        REQUIRE(migration1->dsthost() != decision.hosts.at(1));
        REQUIRE_EQ(migration1->dsthost(), c.mainHost);
    } else {
        REQUIRE(migration0 == nullptr);
        auto migration1 = c.f.sch.checkForMigrationOpportunities(*req->mutable_messages(1), decision.groupId);
        REQUIRE(migration1 == nullptr);
    }
    c.finish(req, { c.mainHost, mustMigrate ? c.mainHost : OTHER });
}

void mpiWorld(bool mustMigrate)
{
    MigrationCase c(2, 2);
    faabric::mpi::getMpiWorldRegistry().clear();
    const int worldId = 123, worldSize = 4;
    auto req = faabric::util::batchExecFactory("mpi", "sleep", 1);
    auto* first = req->mutable_messages(0);
    first->set_ismpi(true);
    first->set_mpiworldsize(worldSize);
    first->set_mpiworldid(worldId);
    auto decision = c.f.plannerCli.callFunctions(req);
    REQUIRE_EQ(decision.hosts.at(0), c.mainHost);
    // rank 0 creates the world: three more ranks, two of them on the other host
    first->set_groupid(decision.groupId);
    faabric::mpi::MpiWorld world;
    world.create(*first, worldId, worldSize);
    const int appId = first->appid();
    const int groupId = first->groupid(); // (assigned when the world grew)
    std::vector<std::string> rankHosts;
    for (int r = 0; r < worldSize; r++) {
        rankHosts.push_back(world.getHostForRank(r));
    }
    REQUIRE_EQ(rankHosts, (std::vector<std::string>{ c.mainHost, c.mainHost, OTHER, OTHER }));
    if (mustMigrate) {
        c.updateLocalResources(4, 2);
    }
    auto migration0 = c.f.sch.checkForMigrationOpportunities(*first);
    int newGroupId = groupId;
    if (mustMigrate) {
        REQUIRE(migration0 != nullptr);
        REQUIRE_EQ(migration0->appid(), appId);
        REQUIRE(migration0->groupid() != groupId);
        REQUIRE_EQ(migration0->groupidx(), 0);
        REQUIRE_EQ(migration0->srchost(), migration0->dsthost());
        REQUIRE_EQ(migration0->dsthost(), c.mainHost);
        newGroupId = migration0->groupid();
    } else {
        REQUIRE(migration0 == nullptr);
    }
    // the other ranks' messages were made by the world: only ids matter here
    for (int i = 1; i < worldSize; i++) {
        faabric::Message msg;
        msg.set_appid(appId);
        msg.set_groupid(groupId);
        msg.set_groupidx(i);
        auto migration = c.f.sch.checkForMigrationOpportunities(msg, newGroupId);
        if (mustMigrate) {
            REQUIRE(migration != nullptr);
            REQUIRE_EQ(migration->appid(), appId);
            REQUIRE(migration->groupid() != groupId);
            REQUIRE_EQ(migration->groupidx(), i);
            // everybody ends up next to rank 0
            REQUIRE_EQ(migration->dsthost(), c.mainHost);
        } else {
            REQUIRE(migration == nullptr);
        }
    }
    c.release->store(true);
    world.destroy();
    faabric::mpi::getMpiWorldRegistry().clear();
}
}

TEST_CASE("migration case: two functions, a slot frees up next to the first: the second must move", "[scheduler][migration][cases]")
{
    twoFunctions(true);
}

TEST_CASE("migration case: two functions, nothing changed: nobody moves", "[scheduler][migration][cases]")
{
    twoFunctions(false);
}

TEST_CASE("migration case: an MPI world split over two hosts is gathered onto the first", "[scheduler][migration][cases]")
{
    mpiWorld(true);
}

TEST_CASE("migration case: an MPI world with no better placement stays", "[scheduler][migration][cases]")
{
    mpiWorld(false);
}
