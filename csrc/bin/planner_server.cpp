// Planner daemon.  Three services share the process: the RPC server workers
// talk to, a snapshot server (frozen / migrating apps park their memory images
// here) and the JSON-over-HTTP control API, which runs in the foreground until
// SIGINT / SIGTERM.  (Counterpart of the reference's src/planner/planner_server.cpp.)
#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerEndpointHandler.h>
#include <faabric/planner/PlannerServer.h>
#include <faabric/snapshot/SnapshotServer.h>
#include <faabric/util/config.h>
#include <faabric/util/crash.h>
#include <faabric/util/logging.h>

namespace {
// Starts a background server now and stops it when the scope unwinds
template<typename Server>
struct Running
{
    Server server;
    const char* name;

    explicit Running(const char* nameIn)
      : name(nameIn)
    {
        SPDLOG_INFO("Planner: starting {}", name);
        server.start();
    }

    ~Running()
    {
        SPDLOG_INFO("Planner: stopping {}", name);
        server.stop();
    }
};
}

int main()
{
    faabric::util::initLogging();
    faabric::util::setUpCrashHandler();
    faabric::util::exitWithParentIfAsked();

    Running<faabric::planner::PlannerServer> rpc("RPC server");
    Running<faabric::snapshot::SnapshotServer> snapshots("snapshot server");

    const int httpPort = faabric::util::getSystemConfig().plannerPort;
    const int httpThreads = faabric::planner::getPlanner().getConfig().numthreadshttpserver();
    faabric::endpoint::FaabricEndpoint http(
      httpPort, httpThreads, std::make_shared<faabric::planner::PlannerEndpointHandler>());
    SPDLOG_INFO("Planner: serving HTTP on {} ({} threads)", httpPort, httpThreads);
    http.start(faabric::endpoint::EndpointMode::SIGNAL);
    return 0;
}
