// Scheduling decisions, the policy interface and the three policies.
//
// One header per module: the per-class headers of the reference's layout
// (faabric/batch-scheduler/*.h) forward here, so either include style works.
#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/batch.h>

#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

// ==========================================================================
// batch-scheduler/SchedulingDecision.h
// ==========================================================================
// Placement of the messages of one app onto hosts (GPUs).  Parallel vectors,
// one entry per message (reference: SchedulingDecision.h:58-119).



namespace faabric::batch_scheduler {

// Legacy placement hints and migration strategies: the planner's policies
// (bin-pack / compact / spot) replaced them, the names stay for embedders that
// still pass them around (reference: include/faabric/batch-scheduler/
// SchedulingDecision.h:9-56)
enum SchedulingTopologyHint
{
    NONE,
    CACHED,
    FORCE_LOCAL,
    NEVER_ALONE,
    UNDERFULL,
};

const std::unordered_map<std::string, SchedulingTopologyHint> strToTopologyHint = {
    { "NONE", SchedulingTopologyHint::NONE },
    { "CACHED", SchedulingTopologyHint::CACHED },
    { "FORCE_LOCAL", SchedulingTopologyHint::FORCE_LOCAL },
    { "NEVER_ALONE", SchedulingTopologyHint::NEVER_ALONE },
    { "UNDERFULL", SchedulingTopologyHint::UNDERFULL },
};

const std::unordered_map<SchedulingTopologyHint, std::string> topologyHintToStr = {
    { SchedulingTopologyHint::NONE, "NONE" },
    { SchedulingTopologyHint::CACHED, "CACHED" },
    { SchedulingTopologyHint::FORCE_LOCAL, "FORCE_LOCAL" },
    { SchedulingTopologyHint::NEVER_ALONE, "NEVER_ALONE" },
    { SchedulingTopologyHint::UNDERFULL, "UNDERFULL" },
};

enum MigrationStrategy
{
    BIN_PACK,
    EMPTY_HOSTS
};

class SchedulingDecision
{
  public:
    static SchedulingDecision fromPointToPointMappings(
      faabric::PointToPointMappings& mappings);

    SchedulingDecision(uint32_t appIdIn, int32_t groupIdIn);

    bool operator==(const SchedulingDecision& rhs) const = default;

    uint32_t appId = 0;

    int32_t groupId = 0;

    int32_t nFunctions = 0;

    std::vector<std::string> hosts;

    std::vector<int32_t> messageIds;

    std::vector<int32_t> appIdxs;

    std::vector<int32_t> groupIdxs;

    // "MPI port" of each message: on GPUs the mailbox / stream slot index
    std::vector<int32_t> mpiPorts;

    std::string returnHost;

    // True when every message sits on one host AND that host is this one
    bool isSingleHost() const;

    void addMessage(const std::string& host, const faabric::Message& msg);

    void addMessage(const std::string& host, int32_t messageId, int32_t appIdx);

    void addMessage(const std::string& host,
                    int32_t messageId,
                    int32_t appIdx,
                    int32_t groupIdx);

    void addMessageInPosition(int32_t pos,
                              const std::string& host,
                              int32_t messageId,
                              int32_t appIdx,
                              int32_t groupIdx,
                              int32_t mpiPort);

    // Returns the MPI port the message vacated
    int32_t removeMessage(int32_t messageId);

    std::set<std::string> uniqueHosts();

    void print(const std::string& logLevel = "debug");

    std::string toString() const;
};

}

// ==========================================================================
// batch-scheduler/BatchScheduler.h
// ==========================================================================
// Batch scheduling policies used by the planner: bin-pack, compact, spot.
// One greedy packing engine; a policy is three hooks (which hosts are
// eligible, how they are ordered for each decision type, and whether a
// re-distribution is worth migrating for).  Behaviour matches the reference
// policies (src/batch-scheduler/{BinPack,Compact,Spot}Scheduler.cpp).



#define DO_NOT_MIGRATE -98
#define DO_NOT_MIGRATE_DECISION                                                \
    faabric::batch_scheduler::SchedulingDecision(DO_NOT_MIGRATE, DO_NOT_MIGRATE)
#define NOT_ENOUGH_SLOTS -99
#define NOT_ENOUGH_SLOTS_DECISION                                              \
    faabric::batch_scheduler::SchedulingDecision(NOT_ENOUGH_SLOTS,             \
                                                 NOT_ENOUGH_SLOTS)
#define MUST_FREEZE -97
#define MUST_FREEZE_DECISION                                                   \
    faabric::batch_scheduler::SchedulingDecision(MUST_FREEZE, MUST_FREEZE)

// Hosts tainted with this address are being evicted (spot policy)
#define MUST_EVICT_IP "E.VI.CT.ME"

namespace faabric::batch_scheduler {

typedef std::pair<std::shared_ptr<BatchExecuteRequest>,
                  std::shared_ptr<SchedulingDecision>>
  InFlightPair;

typedef std::map<int32_t, InFlightPair> InFlightReqs;

struct HostState
{
    HostState(const std::string& ipIn, int slotsIn, int usedSlotsIn)
      : ip(ipIn)
      , slots(slotsIn)
      , usedSlots(usedSlotsIn)
    {}

    std::string ip;
    int slots;
    int usedSlots;
};
typedef std::shared_ptr<HostState> Host;
typedef std::map<std::string, Host> HostMap;

// NEW          first time the app is scheduled
// DIST_CHANGE  in-flight app asking to be re-distributed (MIGRATION request)
// SCALE_CHANGE in-flight app adding messages (the request holds only the NEW
//              messages, not the total)
enum DecisionType
{
    NO_DECISION_TYPE = 0,
    NEW = 1,
    DIST_CHANGE = 2,
    SCALE_CHANGE = 3,
};

class BatchScheduler
{
  public:
    virtual ~BatchScheduler() = default;

    static DecisionType getDecisionType(
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    virtual std::shared_ptr<SchedulingDecision> makeSchedulingDecision(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req) = 0;

  protected:
    static int numSlots(const Host& host) { return host->slots; }

    static int numUsedSlots(const Host& host) { return host->usedSlots; }

    static int numSlotsAvailable(const Host& host)
    {
        return std::max<int>(0, numSlots(host) - numUsedSlots(host));
    }

    static void claimSlots(Host& host, int numSlotsToClaim)
    {
        host->usedSlots =
          std::min<int>(numSlots(host), host->usedSlots + numSlotsToClaim);
    }

    static void freeSlots(Host& host, int numSlotsToFree)
    {
        host->usedSlots = std::max<int>(0, host->usedSlots - numSlotsToFree);
    }

    static std::string getIp(const Host& host) { return host->ip; }

    virtual bool isFirstDecisionBetter(
      std::shared_ptr<SchedulingDecision> decisionA,
      std::shared_ptr<SchedulingDecision> decisionB) = 0;

    virtual std::vector<Host> getSortedHosts(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      const DecisionType& decisionType) = 0;
};

// Shared greedy engine.  Subclasses customise through the virtual hooks.
class GreedyPackScheduler : public BatchScheduler
{
  public:
    std::shared_ptr<SchedulingDecision> makeSchedulingDecision(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req) override;

  protected:
    // How hosts are ordered for a DIST_CHANGE decision once the app's own
    // slots have been handed back
    enum class MigrationOrder
    {
        MostFreeThenAppFrequency, // bin-pack
        FullestFirst,             // compact
        AppFrequencyFirst         // spot
    };

    // Removes ineligible hosts from the map; returns the removed addresses
    virtual std::set<std::string> filterHosts(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    virtual MigrationOrder migrationOrder() const = 0;

    // Decide what a DIST_CHANGE request resolves to given the fresh packing
    virtual std::shared_ptr<SchedulingDecision> resolveDistChange(
      HostMap& hostMap,
      const std::set<std::string>& removedHosts,
      std::shared_ptr<SchedulingDecision> fresh,
      std::shared_ptr<SchedulingDecision> old,
      int numLeftToSchedule) = 0;

    // Restrict OpenMP single-host hinted apps to the best host
    virtual bool honourSingleHostHint() const { return false; }

    std::vector<Host> getSortedHosts(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      const DecisionType& decisionType) override;

    bool isFirstDecisionBetter(
      std::shared_ptr<SchedulingDecision> decisionA,
      std::shared_ptr<SchedulingDecision> decisionB) override;

    // Keeps the host histogram of `fresh` while moving as few messages of
    // `old` as possible
    static std::shared_ptr<SchedulingDecision> minimiseNumOfMigrations(
      std::shared_ptr<SchedulingDecision> fresh,
      std::shared_ptr<SchedulingDecision> old);

    static std::map<std::string, int> hostHistogram(
      const std::shared_ptr<SchedulingDecision>& decision);
};

// Fewest hosts, then fewest cross-host links
class BinPackScheduler final : public GreedyPackScheduler
{
  protected:
    MigrationOrder migrationOrder() const override
    {
        return MigrationOrder::MostFreeThenAppFrequency;
    }
    bool honourSingleHostHint() const override { return true; }
    std::shared_ptr<SchedulingDecision> resolveDistChange(
      HostMap& hostMap,
      const std::set<std::string>& removedHosts,
      std::shared_ptr<SchedulingDecision> fresh,
      std::shared_ptr<SchedulingDecision> old,
      int numLeftToSchedule) override;
};

// Multi-tenant: never share a host with another tenant (BER subType), and
// migrate when that empties more hosts
class CompactScheduler final : public GreedyPackScheduler
{
  public:
    // Compact compares decisions through the host map, not pairwise
    bool isFirstDecisionBetter(HostMap& hostMap,
                               std::shared_ptr<SchedulingDecision> newDecision,
                               std::shared_ptr<SchedulingDecision> oldDecision);

  protected:
    std::set<std::string> filterHosts(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req) override;
    MigrationOrder migrationOrder() const override
    {
        return MigrationOrder::FullestFirst;
    }
    bool isFirstDecisionBetter(
      std::shared_ptr<SchedulingDecision> decisionA,
      std::shared_ptr<SchedulingDecision> decisionB) override;
    std::shared_ptr<SchedulingDecision> resolveDistChange(
      HostMap& hostMap,
      const std::set<std::string>& removedHosts,
      std::shared_ptr<SchedulingDecision> fresh,
      std::shared_ptr<SchedulingDecision> old,
      int numLeftToSchedule) override;
};

// Spot VMs / GPUs being drained: move off evicted hosts, or freeze the app if
// there is nowhere to go
class SpotScheduler final : public GreedyPackScheduler
{
  protected:
    std::set<std::string> filterHosts(
      HostMap& hostMap,
      const InFlightReqs& inFlightReqs,
      std::shared_ptr<faabric::BatchExecuteRequest> req) override;
    MigrationOrder migrationOrder() const override
    {
        return MigrationOrder::AppFrequencyFirst;
    }
    bool isFirstDecisionBetter(
      std::shared_ptr<SchedulingDecision> decisionA,
      std::shared_ptr<SchedulingDecision> decisionB) override;
    std::shared_ptr<SchedulingDecision> resolveDistChange(
      HostMap& hostMap,
      const std::set<std::string>& removedHosts,
      std::shared_ptr<SchedulingDecision> fresh,
      std::shared_ptr<SchedulingDecision> old,
      int numLeftToSchedule) override;
};

std::shared_ptr<BatchScheduler> getBatchScheduler();

void resetBatchScheduler();

void resetBatchScheduler(const std::string& newMode);

}

// ==========================================================================
// batch-scheduler/BinPackScheduler.h
// ==========================================================================

// ==========================================================================
// batch-scheduler/CompactScheduler.h
// ==========================================================================

// ==========================================================================
// batch-scheduler/DecisionCache.h
// ==========================================================================
// Remembers where an app of a given size was placed so a repeat invocation can
// skip scheduling (reference: src/batch-scheduler/DecisionCache.cpp:7-78)



namespace faabric::batch_scheduler {

class CachedDecision
{
  public:
    CachedDecision(const std::vector<std::string>& hostsIn, int groupIdIn);

    std::vector<std::string> getHosts() { return hosts; }

    int getGroupId() const { return groupId; }

  private:
    std::vector<std::string> hosts;
    int groupId = 0;
};

class DecisionCache
{
  public:
    std::shared_ptr<CachedDecision> getCachedDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    void addCachedDecision(std::shared_ptr<faabric::BatchExecuteRequest> req,
                           SchedulingDecision& decision);

    void clear();

  private:
    std::string getCacheKey(std::shared_ptr<faabric::BatchExecuteRequest> req);

    std::shared_mutex mx;
    std::unordered_map<std::string, std::shared_ptr<CachedDecision>>
      cachedDecisions;
};

DecisionCache& getSchedulingDecisionCache();

}

// ==========================================================================
// batch-scheduler/SpotScheduler.h
// ==========================================================================

