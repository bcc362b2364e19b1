// One case per scenario of the reference's two-host message-count tables
// (reference: tests/test/mpi/test_remote_mpi_worlds.cpp:34-339): a 4-rank
// world, ranks 0,1 here and 2,3 on another host, driven through MpiWorld in
// mock mode.  Each row: who calls, who is the root, how many messages the
// caller must emit, to whom, and with which element counts.
#include "fixtures.h"

#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/mpi/MpiWorld.h>
#include <faabric/mpi/mpi.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/util/config.h>
#include <faabric/util/func.h>
#include <faabric/util/testing.h>

#include <set>

using namespace tests;
using faabric::mpi::MpiMessageType;

namespace {
struct RemoteWorlds
{
    static constexpr int worldId = 5151;
    static constexpr int groupId = 6161;
    static constexpr int worldSize = 4;
    std::string thisHost = faabric::util::getSystemConfig().endpointHost;
    std::string otherHost = "192.0.2.77";
    faabric::Message msg = faabric::util::messageFactory("mpi", "remote-cases");
    faabric::mpi::MpiWorld here;
    faabric::mpi::MpiWorld there;

    RemoteWorlds()
    {
        faabric::util::setMockMode(true);
        faabric::mpi::clearMpiMockedMessages();
        msg.set_ismpi(true);
        msg.set_mpiworldid(worldId);
        msg.set_mpiworldsize(worldSize);
        msg.set_groupid(groupId);
        faabric::batch_scheduler::SchedulingDecision decision(msg.appid(), groupId);
        for (int r = 0; r < worldSize; r++) {
            decision.addMessage(r < 2 ? thisHost : otherHost, msg.id() + r, r, r);
        }
        faabric::transport::getPointToPointBroker().setUpLocalMappingsFromSchedulingDecision(decision);
        here.initialiseFromMsg(msg);
        there.overrideHost(otherHost);
        there.initialiseFromMsg(msg);
    }

    ~RemoteWorlds()
    {
        faabric::mpi::clearMpiMockedMessages();
        faabric::transport::getPointToPointBroker().clear();
        faabric::util::setMockMode(false);
    }

    faabric::mpi::MpiWorld& worldOf(int rank) { return rank < 2 ? here : there; }
};

struct Sent
{
    size_t n;
    std::set<int> receivers;
    std::set<int> counts;
};

Sent sentBy(int rank)
{
    Sent s{ 0, {}, {} };
    for (const auto& m : faabric::mpi::getMpiMockedMessages(rank)) {
        s.n++;
        s.receivers.insert(m.recvRank);
        s.counts.insert(m.count);
    }
    return s;
}

void broadcastRow(int root, int caller, size_t nMsgs, std::set<int> receivers)
{
    RemoteWorlds w;
    std::vector<int> data = { 0, 1, 2 };
    w.worldOf(caller).broadcast(root, caller, BYTES(data.data()), MPI_INT, (int)data.size(), MpiMessageType::BROADCAST);
    Sent s = sentBy(caller);
    REQUIRE_EQ(s.n, nMsgs);
    REQUIRE(s.receivers == receivers);
    if (nMsgs > 0) {
        REQUIRE(s.counts == std::set<int>{ 3 });
    }
}

void reduceRow(int caller, int root, size_t nMsgs, std::set<int> receivers)
{
    RemoteWorlds w;
    std::vector<int> data = { 0, 1, 2 };
    std::vector<int> out(data.size(), -1);
    w.worldOf(caller).reduce(caller, root, BYTES(data.data()), BYTES(out.data()), MPI_INT, (int)data.size(), MPI_SUM);
    Sent s = sentBy(caller);
    REQUIRE_EQ(s.n, nMsgs);
    REQUIRE(s.receivers == receivers);
    // the caller's send buffer is never the accumulator
    REQUIRE(data[0] == 0 && data[1] == 1 && data[2] == 2);
}

void gatherRow(int caller, int root, size_t nMsgs, std::set<int> receivers, std::set<int> counts)
{
    RemoteWorlds w;
    std::vector<int> data = { 0, 1, 2 };
    std::vector<int> out(RemoteWorlds::worldSize * data.size(), -1);
    w.worldOf(caller).gather(caller, root, BYTES(data.data()), MPI_INT, 3, BYTES(out.data()), MPI_INT, 3);
    Sent s = sentBy(caller);
    REQUIRE_EQ(s.n, nMsgs);
    REQUIRE(s.receivers == receivers);
    REQUIRE(s.counts == counts);
}
}

// ---- broadcast: root feeds its co-located ranks and ONE leader per other host
TEST_CASE("mpi remote case: broadcast, the root is its host's leader", "[mpi][mock][cases]")
{
    broadcastRow(0, 0, 2, { 1, 2 });
}

TEST_CASE("mpi remote case: broadcast, the root is not its host's leader", "[mpi][mock][cases]")
{
    broadcastRow(1, 1, 2, { 0, 2 });
}

TEST_CASE("mpi remote case: broadcast, a leader co-located with the root forwards nothing", "[mpi][mock][cases]")
{
    broadcastRow(1, 0, 0, {});
}

TEST_CASE("mpi remote case: broadcast, a non-leader co-located with the root forwards nothing", "[mpi][mock][cases]")
{
    broadcastRow(0, 1, 0, {});
}

TEST_CASE("mpi remote case: broadcast, the leader of the other host forwards to its host", "[mpi][mock][cases]")
{
    broadcastRow(0, 2, 1, { 3 });
}

TEST_CASE("mpi remote case: broadcast, a leaf of the other host forwards nothing", "[mpi][mock][cases]")
{
    broadcastRow(0, 3, 0, {});
}

// ---- reduce: everyone hands its data to the root or to its own host's leader
TEST_CASE("mpi remote case: reduce, the root (a leader) sends nothing", "[mpi][mock][cases]")
{
    reduceRow(0, 0, 0, {});
}

TEST_CASE("mpi remote case: reduce, the root (not a leader) sends nothing", "[mpi][mock][cases]")
{
    reduceRow(1, 1, 0, {});
}

TEST_CASE("mpi remote case: reduce, the leader co-located with the root sends to the root", "[mpi][mock][cases]")
{
    reduceRow(0, 1, 1, { 1 });
}

TEST_CASE("mpi remote case: reduce, a rank co-located with the root sends to the root", "[mpi][mock][cases]")
{
    reduceRow(1, 0, 1, { 0 });
}

TEST_CASE("mpi remote case: reduce, the other host's leader sends one message to the root", "[mpi][mock][cases]")
{
    reduceRow(2, 0, 1, { 0 });
}

TEST_CASE("mpi remote case: reduce, a leaf of the other host sends to its leader", "[mpi][mock][cases]")
{
    reduceRow(3, 0, 1, { 2 });
}

// ---- gather: like reduce, and the remote leader packs its host's chunks
TEST_CASE("mpi remote case: gather, the root (a leader) sends nothing", "[mpi][mock][cases]")
{
    gatherRow(0, 0, 0, {}, {});
}

TEST_CASE("mpi remote case: gather, the root (not a leader) sends nothing", "[mpi][mock][cases]")
{
    gatherRow(1, 1, 0, {}, {});
}

TEST_CASE("mpi remote case: gather, the leader co-located with the root sends its chunk", "[mpi][mock][cases]")
{
    gatherRow(0, 1, 1, { 1 }, { 3 });
}

TEST_CASE("mpi remote case: gather, a rank co-located with the root sends its chunk", "[mpi][mock][cases]")
{
    gatherRow(1, 0, 1, { 0 }, { 3 });
}

TEST_CASE("mpi remote case: gather, the other host's leader sends both chunks in one message", "[mpi][mock][cases]")
{
    gatherRow(2, 0, 1, { 0 }, { 6 });
}

TEST_CASE("mpi remote case: gather, a leaf of the other host sends its chunk to its leader", "[mpi][mock][cases]")
{
    gatherRow(3, 0, 1, { 2 }, { 3 });
}
