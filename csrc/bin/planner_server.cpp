// Planner daemon: RPC server + snapshot server + JSON-over-HTTP control API
// (reference: src/planner/planner_server.cpp)
#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerEndpointHandler.h>
#include <faabric/planner/PlannerServer.h>
#include <faabric/snapshot/SnapshotServer.h>
#include <faabric/util/config.h>
#include <faabric/util/crash.h>
#include <faabric/util/logging.h>

int main()
{
    faabric::util::initLogging();
    faabric::util::setUpCrashHandler();

    SPDLOG_INFO("Starting planner server");
    faabric::planner::PlannerServer plannerServer;
    plannerServer.start();

    // Snapshots of frozen / migrating apps are parked on the planner
    SPDLOG_INFO("Starting planner snapshot server");
    faabric::snapshot::SnapshotServer snapshotServer;
    snapshotServer.start();

    SPDLOG_INFO("Starting planner endpoint");
    faabric::endpoint::FaabricEndpoint endpoint(faabric::util::getSystemConfig().plannerPort,
                                                faabric::planner::getPlanner().getConfig().numthreadshttpserver(),
                                                std::make_shared<faabric::planner::PlannerEndpointHandler>());
    // Blocks until SIGINT / SIGTERM
    endpoint.start(faabric::endpoint::EndpointMode::SIGNAL);

    SPDLOG_INFO("Planner snapshot server shutting down");
    snapshotServer.stop();
    SPDLOG_INFO("Planner server shutting down");
    plannerServer.stop();
    return 0;
}
