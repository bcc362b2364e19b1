// One case per case of the reference's point-to-point tests
// (reference: tests/test/transport/test_point_to_point.cpp:24-520,
// test_point_to_point_groups.cpp:40-470)
#include "harness.h"

#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/transport/PointToPointClient.h>
#include <faabric/transport/PointToPointServer.h>
#include <faabric/util/config.h>
#include <faabric/util/func.h>
#include <faabric/util/gids.h>
#include <faabric/util/testing.h>

#include <atomic>
#include <thread>

using namespace faabric::transport;

namespace {
struct PtpCase
{
    PointToPointBroker& broker = getPointToPointBroker();
    PointToPointServer server;
    std::string thisHost = faabric::util::getSystemConfig().endpointHost;

    PtpCase()
    {
        faabric::util::setMockMode(false);
        broker.clear();
        clearPointToPointClients();
        server.start();
    }

    ~PtpCase()
    {
        server.stop();
        broker.clear();
        faabric::util::setMockMode(false);
    }

    // `n` members, all on this host
    std::shared_ptr<PointToPointGroup> localGroup(int appId, int groupId, int n)
    {
        faabric::batch_scheduler::SchedulingDecision d(appId, groupId);
        for (int i = 0; i < n; i++) {
            d.addMessage(thisHost, faabric::util::generateGid(), i, i);
        }
        broker.setUpLocalMappingsFromSchedulingDecision(d);
        return PointToPointGroup::getGroup(groupId);
    }
};
}

TEST_CASE("ptp case: mappings sent by a client register idxs and hosts per group", "[transport][ptp][cases]")
{
    PtpCase f;
    const int groupA = 321, groupB = 543;
    REQUIRE(f.broker.getIdxsRegisteredForGroup(groupA).empty());
    REQUIRE(f.broker.getHostsRegisteredForGroup(groupA).empty());
    faabric::PointToPointMappings a, b;
    a.set_appid(123);
    a.set_groupid(groupA);
    b.set_appid(345);
    b.set_groupid(groupB);
    // (idxs overlap between the groups on purpose)
    auto* a1 = a.add_mappings();
    a1->set_appidx(1);
    a1->set_groupidx(3);
    a1->set_host("host-a");
    auto* a2 = a.add_mappings();
    a2->set_appidx(2);
    a2->set_groupidx(4);
    a2->set_host("host-b");
    auto* b1 = b.add_mappings();
    b1->set_appidx(1);
    b1->set_groupidx(3);
    b1->set_host("host-a");
    PointToPointClient cli(f.thisHost);
    cli.sendMappings(a);
    cli.sendMappings(b);
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(groupA).size(), 2u);
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(groupB).size(), 1u);
    REQUIRE_EQ(f.broker.getHostsRegisteredForGroup(groupA).size(), 2u);
    REQUIRE_EQ(f.broker.getHostsRegisteredForGroup(groupB).size(), 1u);
    REQUIRE_EQ(f.broker.getHostForReceiver(groupA, 3), std::string("host-a"));
    REQUIRE_EQ(f.broker.getHostForReceiver(groupA, 4), std::string("host-b"));
    REQUIRE_EQ(f.broker.getHostForReceiver(groupB, 3), std::string("host-a"));
}

TEST_CASE("ptp case: send and receive between two members, both ways", "[transport][ptp][cases]")
{
    PtpCase f;
    const int groupId = 111;
    f.localGroup(11, groupId, 2);
    std::vector<uint8_t> sentA = { 0, 1, 2, 3 }, sentB = { 3, 4, 5 };
    std::vector<uint8_t> gotByB;
    std::thread peer([&] {
        gotByB = f.broker.recvMessage(groupId, 0, 1);
        f.broker.sendMessage(groupId, 1, 0, sentB.data(), sentB.size());
        f.broker.resetThreadLocalCache();
    });
    f.broker.sendMessage(groupId, 0, 1, sentA.data(), sentA.size());
    auto gotByA = f.broker.recvMessage(groupId, 1, 0);
    peer.join();
    REQUIRE(gotByB == sentA);
    REQUIRE(gotByA == sentB);
}

TEST_CASE("ptp case: messages sent with ordering on are received in sending order", "[transport][ptp][cases]")
{
    PtpCase f;
    const int groupId = 112;
    f.localGroup(12, groupId, 3);
    const int n = 100;
    std::thread sender([&] {
        for (int i = 0; i < n; i++) {
            // to two receivers, interleaved
            f.broker.sendMessage(groupId, 0, 1, (const uint8_t*)&i, sizeof(i), true);
            f.broker.sendMessage(groupId, 0, 2, (const uint8_t*)&i, sizeof(i), true);
        }
        f.broker.resetThreadLocalCache();
    });
    for (int r : { 1, 2 }) {
        for (int i = 0; i < n; i++) {
            auto m = f.broker.recvMessage(groupId, 0, r, true);
            REQUIRE_EQ(*(const int*)m.data(), i);
        }
    }
    sender.join();
    // delivered out of order by hand: still handed out in sequence
    int v0 = 0, v1 = 1, v2 = 2;
    f.broker.deliverLocally(groupId, 1, 2, (const uint8_t*)&v2, sizeof(int), 2);
    f.broker.deliverLocally(groupId, 1, 2, (const uint8_t*)&v0, sizeof(int), 0);
    f.broker.deliverLocally(groupId, 1, 2, (const uint8_t*)&v1, sizeof(int), 1);
    for (int want : { 0, 1, 2 }) {
        REQUIRE_EQ(*(const int*)f.broker.recvMessage(groupId, 1, 2, true).data(), want);
    }
}

TEST_CASE("ptp case: a scheduling decision sets the local mappings and sends them to the other hosts", "[transport][ptp][cases]")
{
    PtpCase f;
    faabric::util::setMockMode(true);
    clearSentMessages();
    const int appId = 1, groupId = 113;
    faabric::batch_scheduler::SchedulingDecision d(appId, groupId);
    d.addMessage(f.thisHost, 101, 0, 0);
    d.addMessage("host-a", 102, 1, 1);
    d.addMessage("host-b", 103, 2, 2);
    d.addMessage("host-a", 104, 3, 3);
    f.broker.setAndSendMappingsFromSchedulingDecision(d);
    REQUIRE(f.broker.getHostsRegisteredForGroup(groupId) == (std::set<std::string>{ f.thisHost, "host-a", "host-b" }));
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(groupId).size(), 4u);
    REQUIRE_EQ(f.broker.getHostForReceiver(groupId, 0), f.thisHost);
    REQUIRE_EQ(f.broker.getHostForReceiver(groupId, 1), std::string("host-a"));
    REQUIRE_EQ(f.broker.getHostForReceiver(groupId, 2), std::string("host-b"));
    REQUIRE_EQ(f.broker.getHostForReceiver(groupId, 3), std::string("host-a"));
    // one copy of the whole mapping per remote host
    auto sent = getSentMappings();
    REQUIRE_EQ(sent.size(), 2u);
    std::set<std::string> to;
    for (auto& [host, mappings] : sent) {
        to.insert(host);
        REQUIRE_EQ(mappings.appid(), appId);
        REQUIRE_EQ(mappings.groupid(), groupId);
        REQUIRE_EQ(mappings.mappings_size(), 4);
    }
    REQUIRE(to == (std::set<std::string>{ "host-a", "host-b" }));
}

TEST_CASE("ptp case: waiting for the mappings of a group blocks until they are set, then never again", "[transport][ptp][cases]")
{
    PtpCase f;
    const int appId = 123, groupId = 345;
    std::atomic<int> shared{ 5 };
    faabric::batch_scheduler::SchedulingDecision d(appId, groupId);
    d.addMessage(f.thisHost, 1, 0, 0);
    std::thread enabler([&] {
        std::this_thread::sleep_for(std::chrono::milliseconds(300));
        shared.fetch_add(100);
        f.broker.setUpLocalMappingsFromSchedulingDecision(d);
    });
    f.broker.waitForMappingsOnThisHost(groupId);
    REQUIRE_EQ(shared.load(), 105);
    f.broker.waitForMappingsOnThisHost(groupId); // returns at once
    enabler.join();
}

TEST_CASE("ptp case: distributed lock and unlock through the client, plain and recursive", "[transport][ptp][cases]")
{
    for (bool recursive : { false, true }) {
        PtpCase f;
        const int appId = 999, groupId = 888;
        auto group = f.localGroup(appId, groupId, 2);
        PointToPointClient cli(f.thisHost);
        REQUIRE_EQ(group->getLockOwner(recursive), NO_LOCK_OWNER_IDX);
        cli.groupLock(appId, groupId, 1, recursive);
        // the grant is a message to the locker
        f.broker.recvMessage(groupId, POINT_TO_POINT_MAIN_IDX, 1);
        REQUIRE_EQ(group->getLockOwner(recursive), 1);
        cli.groupUnlock(appId, groupId, 1, recursive);
        for (int i = 0; i < 200 && group->getLockOwner(recursive) != NO_LOCK_OWNER_IDX; i++) {
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
        REQUIRE_EQ(group->getLockOwner(recursive), NO_LOCK_OWNER_IDX);
    }
}

TEST_CASE("ptp case: clearing a group removes its mappings and the group", "[transport][ptp][cases]")
{
    PtpCase f;
    f.localGroup(1, 501, 3);
    f.localGroup(2, 502, 2);
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(501).size(), 3u);
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(502).size(), 2u);
    REQUIRE(PointToPointGroup::groupExists(501));
    f.broker.clearGroup(501);
    REQUIRE(f.broker.getIdxsRegisteredForGroup(501).empty());
    REQUIRE(!PointToPointGroup::groupExists(501));
    // the other group is untouched
    REQUIRE_EQ(f.broker.getIdxsRegisteredForGroup(502).size(), 2u);
    REQUIRE(PointToPointGroup::groupExists(502));
    f.broker.clear();
    REQUIRE(f.broker.getIdxsRegisteredForGroup(502).empty());
    REQUIRE(!PointToPointGroup::groupExists(502));
}

TEST_CASE("ptp case: lock requests of a member whose coordinator is elsewhere go to that host", "[transport][ptp][cases]")
{
    for (bool recursive : { false, true }) {
        PtpCase f;
        faabric::util::setMockMode(true);
        const int appId = 7, groupId = 601;
        faabric::batch_scheduler::SchedulingDecision d(appId, groupId);
        d.addMessage("other-host", 1, 0, 0); // the coordinator (idx 0) lives elsewhere
        d.addMessage(f.thisHost, 2, 1, 1);
        f.broker.setUpLocalMappingsFromSchedulingDecision(d);
        auto group = PointToPointGroup::getGroup(groupId);
        // the request leaves for the coordinator's host; nothing answers in mock
        // mode, so the grant (a message 0 -> 1) is put in place beforehand
        uint8_t grant = 0;
        f.broker.deliverLocally(groupId, POINT_TO_POINT_MAIN_IDX, 1, &grant, 1, -1);
        clearSentMessages();
        group->lock(1, recursive);
        group->unlock(1, recursive);
        auto reqs = getSentLockMessages();
        REQUIRE_EQ(reqs.size(), 2u);
        REQUIRE_EQ(std::get<0>(reqs[0]), std::string("other-host"));
        REQUIRE(std::get<1>(reqs[0]) == PointToPointCall::LOCK_GROUP ||
                std::get<1>(reqs[0]) == PointToPointCall::LOCK_GROUP_RECURSIVE);
        REQUIRE_EQ(std::get<2>(reqs[0]).groupid(), groupId);
        REQUIRE_EQ(std::get<2>(reqs[0]).sendidx(), 1);
        REQUIRE(std::get<1>(reqs[1]) == PointToPointCall::UNLOCK_GROUP ||
                std::get<1>(reqs[1]) == PointToPointCall::UNLOCK_GROUP_RECURSIVE);
    }
}

TEST_CASE("ptp case: locking and unlocking, local, with waiters queueing up", "[transport][ptp][cases]")
{
    PtpCase f;
    const int n = 4;
    auto group = f.localGroup(21, 701, n);
    // plain lock: one owner at a time, the others are served in turn
    group->lock(0, false);
    REQUIRE_EQ(group->getLockOwner(false), 0);
    std::atomic<int> entered{ 0 };
    std::vector<std::thread> waiters;
    for (int i = 1; i < n; i++) {
        waiters.emplace_back([&, i] {
            group->lock(i, false);
            entered++;
            group->unlock(i, false);
            f.broker.resetThreadLocalCache();
        });
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    REQUIRE_EQ(entered.load(), 0);
    group->unlock(0, false);
    for (auto& t : waiters) {
        t.join();
    }
    REQUIRE_EQ(entered.load(), n - 1);
    REQUIRE_EQ(group->getLockOwner(false), NO_LOCK_OWNER_IDX);
    // recursive: the owner may take it again and must release as often
    group->lock(2, true);
    group->lock(2, true);
    REQUIRE_EQ(group->getLockOwner(true), 2);
    group->unlock(2, true);
    REQUIRE_EQ(group->getLockOwner(true), 2);
    group->unlock(2, true);
    REQUIRE_EQ(group->getLockOwner(true), NO_LOCK_OWNER_IDX);
}

TEST_CASE("ptp case: the distributed barrier holds everyone until the last member arrives", "[transport][ptp][cases]")
{
    PtpCase f;
    const int n = 5, rounds = 10;
    auto group = f.localGroup(31, 801, n);
    std::atomic<int> arrived{ 0 };
    std::atomic<bool> early{ false };
    std::vector<std::thread> members;
    for (int i = 0; i < n; i++) {
        members.emplace_back([&, i] {
            for (int r = 0; r < rounds; r++) {
                if (i == n - 1) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(10)); // a straggler
                }
                arrived++;
                group->barrier(i);
                if (arrived.load() < (r + 1) * n) {
                    early = true;
                }
                group->barrier(i);
            }
            f.broker.resetThreadLocalCache();
        });
    }
    for (auto& t : members) {
        t.join();
    }
    REQUIRE(!early.load());
    REQUIRE_EQ(arrived.load(), n * rounds);
}

TEST_CASE("ptp case: local try-lock per group", "[transport][ptp][cases]")
{
    PtpCase f;
    auto group = f.localGroup(11, 111, 5);
    auto other = f.localGroup(22, 222, 5);
    REQUIRE(group->localTryLock());
    REQUIRE(other->localTryLock());
    REQUIRE(!group->localTryLock());
    REQUIRE(!other->localTryLock());
    group->localUnlock();
    REQUIRE(group->localTryLock());
    REQUIRE(!other->localTryLock());
    group->localUnlock();
    other->localUnlock();
    REQUIRE(group->localTryLock());
    REQUIRE(other->localTryLock());
    group->localUnlock();
    other->localUnlock();
}

TEST_CASE("ptp case: notify and await: the main member returns once every other member has notified", "[transport][ptp][cases]")
{
    PtpCase f;
    const int n = 4;
    auto group = f.localGroup(41, 901, n);
    std::atomic<int> notified{ 0 };
    std::vector<std::thread> members;
    for (int i = 1; i < n; i++) {
        members.emplace_back([&, i] {
            std::this_thread::sleep_for(std::chrono::milliseconds(20 * i));
            notified++;
            group->notify(i);
            f.broker.resetThreadLocalCache();
        });
    }
    group->notify(POINT_TO_POINT_MAIN_IDX); // blocks for the n - 1 others
    REQUIRE_EQ(notified.load(), n - 1);
    for (auto& t : members) {
        t.join();
    }
}
