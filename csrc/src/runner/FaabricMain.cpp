#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/FaabricMain.h>
#include <faabric/util/crash.h>
#include <faabric/util/hwloc.h>
#include <faabric/state/InMemoryStateRegistry.h>
#include <faabric/state/State.h>
#include <faabric/util/logging.h>
#include <faabric/util/timing.h>

#include <cuda_runtime.h>

namespace faabric::runner {

FaabricMain::FaabricMain(std::shared_ptr<faabric::executor::ExecutorFactory> execFactory)
  : stateServer(faabric::state::getGlobalState())
{
    faabric::executor::setExecutorFactory(std::move(execFactory));
}

// Bind the worker to its GPU before any server can hand out device work.
// Not having a GPU is fine: the runtime then runs its host paths only.
static void bindDevice()
{
    const auto& conf = faabric::util::getSystemConfig();
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        SPDLOG_INFO("No CUDA device visible, running host-only");
        return;
    }
    // FAABRIC_GPUS is a comma-separated list; a worker owns the first entry
    // for its default context (MPI ranks pick theirs from the decision)
    int dev = conf.gpus.empty() ? 0 : std::atoi(conf.gpus.c_str());
    if (dev < 0 || dev >= count) {
        SPDLOG_WARN("GPU {} out of range ({} devices), using 0", dev, count);
        dev = 0;
    }
    if (cudaSetDevice(dev) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) {
        cudaGetLastError();
        SPDLOG_WARN("Could not initialise CUDA device {}", dev);
        return;
    }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, dev);
    SPDLOG_INFO("Worker bound to GPU {} ({}, {} SMs, {} MiB)",
                dev,
                prop.name,
                prop.multiProcessorCount,
                (size_t)prop.totalGlobalMem >> 20);
}

void FaabricMain::startBackground()
{
    // Crash handler
    faabric::util::setUpCrashHandler();

    PROF_BEGIN

    bindDevice();

    // Start basics
    startRunner();

    // In-memory state, snapshots and point-to-point messaging
    startStateServer();
    startSnapshotServer();
    startPointToPointServer();

    // Work sharing
    startFunctionCallServer();

    PROF_SUMMARY
}

void FaabricMain::startRunner()
{
    // Ensure we can ping the planner, then make this host available
    faabric::planner::getPlannerClient().ping();
    // The planner answers: let it arbitrate state mains so that every worker
    // process agrees on them
    faabric::state::getInMemoryStateRegistry().setShared(true);
    auto& sch = faabric::scheduler::getScheduler();
    sch.addHostToGlobalSet();
}

void FaabricMain::startFunctionCallServer()
{
    SPDLOG_INFO("Starting function call server");
    functionServer.start();
}

void FaabricMain::startSnapshotServer()
{
    SPDLOG_INFO("Starting snapshot server");
    snapshotServer.start();
}

void FaabricMain::startPointToPointServer()
{
    SPDLOG_INFO("Starting point-to-point server");
    pointToPointServer.start();
}

void FaabricMain::startStateServer()
{
    // Skip state server if not in in-memory mode
    const auto& conf = faabric::util::getSystemConfig();
    if (conf.stateMode != "inmemory") {
        SPDLOG_INFO("Not starting state server in state mode {}", conf.stateMode);
        return;
    }
    SPDLOG_INFO("Starting state server");
    stateServer.start();
}

void FaabricMain::shutdown()
{
    faabric::state::getInMemoryStateRegistry().setShared(false);
    SPDLOG_INFO("Removing from global working set");
    auto& sch = faabric::scheduler::getScheduler();
    sch.shutdown();

    SPDLOG_INFO("Waiting for the state server to finish");
    stateServer.stop();

    SPDLOG_INFO("Waiting for the function server to finish");
    functionServer.stop();

    SPDLOG_INFO("Waiting for the snapshot server to finish");
    snapshotServer.stop();

    SPDLOG_INFO("Waiting for the point-to-point server to finish");
    pointToPointServer.stop();

    SPDLOG_INFO("Faabric pool successfully shut down");
}

}
