#pragma once

#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/planner/PlannerApi.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/util/PeriodicBackgroundThread.h>

#include <future>
#include <map>
#include <shared_mutex>

namespace faabric::planner {

// Re-registers this host with the planner every timeout/2
class KeepAliveThread : public faabric::util::PeriodicBackgroundThread
{
  public:
    void doWork() override;

    // Adds (or replaces) the keep-alive of one host served by this process:
    // this host itself and every per-GPU virtual host it exposes
    void setRequest(std::shared_ptr<RegisterHostRequest> hostReqIn);

    // Returns how many hosts are still being kept alive
    size_t removeRequest(const std::string& hostIp);

    std::shared_mutex keepAliveThreadMx;

  private:
    std::map<std::string, std::shared_ptr<RegisterHostRequest>> hostReqs;
};

struct PlannerCache
{
    std::unordered_map<uint32_t, std::promise<std::shared_ptr<faabric::Message>>>
      plannerResults;

    // Snapshots already pushed to the planner, by key
    std::set<std::string> pushedSnapshots;
};

class PlannerClient final : public faabric::transport::MessageEndpointClient
{
  public:
    PlannerClient();

    explicit PlannerClient(const std::string& plannerIp);

    // ------
    // Util
    // ------
    void ping();

    void clearCache();

    // ------
    // Host membership calls
    // ------
    std::vector<Host> getAvailableHosts();

    // Returns the keep-alive timeout (seconds)
    int registerHost(std::shared_ptr<RegisterHostRequest> req);

    void removeHost(std::shared_ptr<RemoveHostRequest> req);

    // ------
    // Scheduling calls
    // ------
    void setMessageResult(std::shared_ptr<faabric::Message> msg);

    // Called by the FunctionCallServer when the planner notifies a result
    void setMessageResultLocally(std::shared_ptr<faabric::Message> msg);

    faabric::Message getMessageResult(int appId, int msgId, int timeoutMs);

    faabric::Message getMessageResult(const faabric::Message& msg,
                                      int timeoutMs);

    std::shared_ptr<faabric::BatchExecuteRequestStatus> getBatchResults(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    faabric::batch_scheduler::SchedulingDecision callFunctions(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    faabric::batch_scheduler::SchedulingDecision getSchedulingDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    int getNumMigrations();

    std::string stateMain(const std::string& user, const std::string& key, const std::string& host, bool claim, bool drop = false);

    void preloadSchedulingDecision(
      std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> preloadDec);

  private:
    std::mutex plannerCacheMx;
    PlannerCache cache;

    faabric::snapshot::SnapshotRegistry& snapshotRegistry;

    faabric::Message doGetMessageResult(
      std::shared_ptr<faabric::Message> msgPtr,
      int timeoutMs);
};

PlannerClient& getPlannerClient();

}
