#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/LocalCluster.h>
#include <faabric/scheduler/Scheduler.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/state/State.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/transport/common.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>

#include <thread>

namespace faabric::runner {

LocalCluster::LocalCluster(std::shared_ptr<faabric::executor::ExecutorFactory> factory,
                           int nVirtualHosts,
                           int slotsPerHost)
  : stateServer(faabric::state::getGlobalState())
{
    auto& conf = faabric::util::getSystemConfig();
    auto& sch = faabric::scheduler::getScheduler();
    faabric::executor::setExecutorFactory(std::move(factory));

    plannerServer.start();
    faabric::planner::getPlanner().reset();
    functionServer.start();
    snapshotServer.start();
    ptpServer.start();
    stateServer.start();
    sch.reset();
    faabric::planner::getPlannerClient().clearCache();

    if (nVirtualHosts <= 0) {
        faabric::HostResources res;
        res.set_slots(slotsPerHost);
        sch.setThisHostResources(res);
        sch.addHostToGlobalSet();
        hostNames.push_back(sch.getThisHost());
        return;
    }
    for (int i = 0; i < nVirtualHosts; i++) {
        std::string name = faabric::util::gpuHostName(i);
        faabric::transport::registerHostAlias(name, conf.endpointHost);
        auto res = std::make_shared<faabric::HostResources>();
        res->set_slots(slotsPerHost);
        sch.addHostToGlobalSet(name, res);
        hostNames.push_back(name);
    }
}

LocalCluster::~LocalCluster()
{
    auto& sch = faabric::scheduler::getScheduler();
    sch.shutdown();
    stateServer.stop();
    ptpServer.stop();
    snapshotServer.stop();
    functionServer.stop();
    faabric::planner::getPlanner().reset();
    plannerServer.stop();
    faabric::transport::clearHostAliases();
    faabric::transport::getPointToPointBroker().clear();
    faabric::snapshot::getSnapshotRegistry().clear();
    faabric::planner::getPlannerClient().clearCache();
    sch.reset();
}

std::shared_ptr<faabric::BatchExecuteRequestStatus> LocalCluster::awaitBatch(
  std::shared_ptr<faabric::BatchExecuteRequest> req,
  int timeoutMs)
{
    // The planner lives in this process: wait on it instead of polling RPCs
    if (!faabric::planner::getPlanner().waitForAppToFinish(req->appid(), timeoutMs)) {
        throw std::runtime_error("Timed out waiting for app " + std::to_string(req->appid()));
    }
    auto status = faabric::planner::getPlannerClient().getBatchResults(req);
    if (status == nullptr) {
        throw std::runtime_error("No results for app " + std::to_string(req->appid()));
    }
    return status;
}

}
