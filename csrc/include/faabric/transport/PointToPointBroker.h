// Group messaging between the functions of an app: idx -> host mappings pushed
// by the planner, ordered send/recv, distributed locks, barriers, notify.
// Reference: include/faabric/transport/PointToPointBroker.h:26-181,
// src/transport/PointToPointBroker.cpp:79-933.  Hosts are GPUs/worker
// processes of one box; local delivery goes through in-process mailboxes, so
// per-pair FIFO order holds by construction.  Device payloads do not travel
// here (MpiWorld moves them over NVLink); this is the control plane.
#pragma once

#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/transport/PointToPointClient.h>
#include <faabric/util/barrier.h>
#include <faabric/util/config.h>
#include <faabric/util/locks.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <shared_mutex>
#include <stack>
#include <string>
#include <unordered_map>
#include <vector>

#define DEFAULT_DISTRIBUTED_TIMEOUT_MS 30000

#define POINT_TO_POINT_MAIN_IDX 0

#define NO_LOCK_OWNER_IDX -1

namespace faabric::transport {

class PointToPointBroker;

class PointToPointGroup
{
  public:
    static std::shared_ptr<PointToPointGroup> getGroup(int groupId);

    static std::shared_ptr<PointToPointGroup> getOrAwaitGroup(int groupId);

    static bool groupExists(int groupId);

    static void addGroup(int appId, int groupId, int groupSize);

    static void addGroupIfNotExists(int appId, int groupId, int groupSize);

    static void clearGroup(int groupId);

    static void clear();

    PointToPointGroup(int appId, int groupIdIn, int groupSizeIn);

    void lock(int groupIdx, bool recursive);

    void unlock(int groupIdx, bool recursive);

    int getLockOwner(bool recursive);

    void localLock();

    void localUnlock();

    bool localTryLock();

    void barrier(int groupIdx);

    void notify(int groupIdx);

    int getNotifyCount();

  private:
    friend class PointToPointServer;

    faabric::util::SystemConfig& conf;

    int timeoutMs = DEFAULT_DISTRIBUTED_TIMEOUT_MS;

    std::string mainHost;
    int appId = 0;
    int groupId = 0;
    int groupSize = 0;

    std::shared_ptr<faabric::util::Barrier> localBarrier;

    std::mutex mx;

    // Local lock (all group members on this host)
    std::timed_mutex localMx;
    std::recursive_timed_mutex localRecursiveMx;

    // Distributed lock state (lives on the main host)
    std::stack<int> recursiveLockOwners;
    int lockOwnerIdx = NO_LOCK_OWNER_IDX;
    std::queue<int> lockWaiters;

    void notifyLocked(int groupIdx);

    void masterLock(int groupIdx, bool recursive);

    void masterUnlock(int groupIdx, bool recursive);

    bool isSingleHost();
};

class PointToPointBroker
{
  public:
    PointToPointBroker();

    std::string getHostForReceiver(int groupId, int recvIdx);

    int getMpiPortForReceiver(int groupId, int recvIdx);

    std::set<std::string> setUpLocalMappingsFromSchedulingDecision(
      const faabric::batch_scheduler::SchedulingDecision& decision);

    void setAndSendMappingsFromSchedulingDecision(
      const faabric::batch_scheduler::SchedulingDecision& decision);

    void sendMappingsFromSchedulingDecision(
      const faabric::batch_scheduler::SchedulingDecision& decision,
      const std::set<std::string>& hostList);

    void waitForMappingsOnThisHost(int groupId);

    std::set<int> getIdxsRegisteredForGroup(int groupId);

    std::set<std::string> getHostsRegisteredForGroup(int groupId);

    void updateHostForIdx(int groupId, int groupIdx, std::string newHost);

    void sendMessage(int groupId,
                     int sendIdx,
                     int recvIdx,
                     const uint8_t* buffer,
                     size_t bufferSize,
                     std::string hostHint,
                     bool mustOrderMsg = false);

    void sendMessage(int groupId,
                     int sendIdx,
                     int recvIdx,
                     const uint8_t* buffer,
                     size_t bufferSize,
                     bool mustOrderMsg = false,
                     int sequenceNum = NO_SEQUENCE_NUM,
                     std::string hostHint = "");

    std::vector<uint8_t> recvMessage(int groupId,
                                     int sendIdx,
                                     int recvIdx,
                                     bool mustOrderMsg = false);

    void clearGroup(int groupId);

    void clear();

    void resetThreadLocalCache();

    void postMigrationHook(int groupId, int groupIdx);

    // Delivery into the local mailbox of (group, send, recv); used by the
    // server for messages that arrive from other hosts
    void deliverLocally(int groupId,
                        int sendIdx,
                        int recvIdx,
                        const uint8_t* buffer,
                        size_t bufferSize,
                        int sequenceNum);

  private:
    faabric::util::SystemConfig& conf;

    std::shared_mutex brokerMutex;

    std::unordered_map<int, std::set<int>> groupIdIdxsMap;
    std::unordered_map<std::string, std::string> mappings;
    std::unordered_map<std::string, int> mpiPortMappings;

    std::unordered_map<int, std::shared_ptr<faabric::util::FlagWaiter>>
      groupFlags;

    // Sender side sequence counters, keyed by (group, send, recv)
    std::mutex seqMx;
    std::unordered_map<std::string, int> sentMsgCount;

    std::shared_ptr<faabric::util::FlagWaiter> getGroupFlag(int groupId);

    Message doRecvMessage(int groupId, int sendIdx, int recvIdx);

    int getAndIncrementSentMsgCount(int groupId, int sendIdx, int recvIdx);
};

PointToPointBroker& getPointToPointBroker();

}
