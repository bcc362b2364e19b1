// The planner: host membership, batch scheduling, result storage, migration
// and freeze/thaw (reference: include/faabric/planner/Planner.h:23-145,
// src/planner/Planner.cpp).  Hosts are GPU workers of the box.
#pragma once

#include <condition_variable>
#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/planner/PlannerState.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotRegistry.h>

#include <shared_mutex>

namespace faabric::planner {

enum FlushType
{
    NoFlushType = 0,
    Hosts = 1,
    Executors = 2,
    SchedulingState = 3,
};

class Planner
{
  public:
    Planner();

    // ----------
    // Planner config
    // ----------
    PlannerConfig getConfig();

    // Seconds without a keep-alive after which a host is dropped
    void setHostKeepAliveTimeout(int seconds);

    void printConfig() const;

    std::string getPolicy();

    void setPolicy(const std::string& newPolicy);

    // ----------
    // Util public API
    // ----------
    bool reset();

    bool flush(faabric::planner::FlushType flushType);

    // ----------
    // Host membership public API
    // ----------
    std::vector<std::shared_ptr<Host>> getAvailableHosts();

    bool registerHost(const Host& hostIn, bool overwrite);

    // Best effort
    void removeHost(const Host& hostIn);

    // ----------
    // Request scheduling public API
    // ----------
    void setMessageResult(std::shared_ptr<faabric::Message> msg);

    // Non-blocking: nullptr if not ready (and the caller is registered as a
    // waiter when it named its main host)
    std::shared_ptr<faabric::Message> getMessageResult(
      std::shared_ptr<faabric::Message> msg);

    void preloadSchedulingDecision(
      int appId,
      std::shared_ptr<batch_scheduler::SchedulingDecision> decision);

    std::shared_ptr<faabric::BatchExecuteRequestStatus> getBatchResults(
      int32_t appId);

    std::shared_ptr<faabric::batch_scheduler::SchedulingDecision>
    getSchedulingDecision(std::shared_ptr<BatchExecuteRequest> req);

    faabric::batch_scheduler::InFlightReqs getInFlightReqs();

    // Blocks until the app has no message in flight (false on timeout).
    // For callers living in the planner's process.
    bool waitForAppToFinish(int32_t appId, int timeoutMs);

    int getNumMigrations();

    std::set<std::string> getNextEvictedHostIps();

    std::map<int32_t, std::shared_ptr<BatchExecuteRequest>> getEvictedReqs();

    // The main entry point: schedule + dispatch
    std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> callBatch(
      std::shared_ptr<BatchExecuteRequest> req);

    // Spot policy: which hosts go away next
    void setNextEvictedVm(const std::set<std::string>& vmIps);

    // State main election: returns the main host of user/key, electing `host`
    // if there is none and `claim` is set ("" = none); `drop` forgets it
    std::string stateMain(const std::string& user, const std::string& key, const std::string& host, bool claim, bool drop);

  private:
    std::shared_mutex plannerMx;
    std::condition_variable_any appFinishedCv;

    void compactInFlightLocked();

    PlannerState state;
    PlannerConfig config;

    faabric::snapshot::SnapshotRegistry& snapshotRegistry;

    // ----------
    // Util private API
    // ----------
    void flushHosts();

    void flushExecutors();

    void flushSchedulingState();

    // ----------
    // Host membership private API
    // ----------
    bool isHostExpired(std::shared_ptr<Host> host, long epochTimeMs = 0);

    // ----------
    // Request scheduling private API
    // ----------
    std::shared_ptr<batch_scheduler::SchedulingDecision>
    getPreloadedSchedulingDecision(
      int32_t appId,
      std::shared_ptr<BatchExecuteRequest> ber);

    void dispatchSchedulingDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> decision);
};

Planner& getPlanner();

}
