// Worker bootstrap: bind the GPU, make sure the planner answers, register this
// host, then bring the four servers up (and down again in reverse order).
// Counterpart of the reference's src/runner/FaabricMain.cpp:11-109.
#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/FaabricMain.h>
#include <faabric/state/InMemoryStateRegistry.h>
#include <faabric/state/State.h>
#include <faabric/util/crash.h>
#include <faabric/util/hwloc.h>
#include <faabric/util/logging.h>
#include <faabric/util/timing.h>

#include <cuda_runtime.h>

namespace faabric::runner {

FaabricMain::FaabricMain(std::shared_ptr<faabric::executor::ExecutorFactory> execFactory)
  : stateServer(faabric::state::getGlobalState())
{
    faabric::executor::setExecutorFactory(std::move(execFactory));
}

// A worker without a visible GPU is fine: it then runs its host paths only.
static void bindDefaultDevice()
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        SPDLOG_INFO("No CUDA device visible, running host-only");
        return;
    }
    // FAABRIC_GPUS is a comma-separated list; the first entry hosts the
    // worker's default context (MPI ranks pick theirs from the decision)
    const std::string& list = faabric::util::getSystemConfig().gpus;
    int dev = list.empty() ? 0 : std::atoi(list.c_str());
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
    faabric::util::setUpCrashHandler();
    PROF_BEGIN
    bindDefaultDevice();
    // Fail fast if there is no planner to talk to
    faabric::planner::getPlannerClient().ping();
    // Everything must answer BEFORE the planner learns about this host: it
    // may dispatch work the instant the registration lands
    restoreCheckpoint();
    startStateServer();
    startSnapshotServer();
    startPointToPointServer();
    startFunctionCallServer();
    startRunner();
    PROF_SUMMARY
}

// FAABRIC_CHECKPOINT_DIR: snapshots (frozen / migrating apps' memory images,
// thread snapshots) survive a worker restart.  The reference has no on-disk
// persistence.
void FaabricMain::restoreCheckpoint()
{
    const std::string& dir = faabric::util::getSystemConfig().checkpointDir;
    if (dir.empty()) {
        return;
    }
    try {
        size_t n = faabric::snapshot::getSnapshotRegistry().restoreFromDir(dir);
        SPDLOG_INFO("Restored {} snapshots from {}", n, dir);
    } catch (const std::exception& e) {
        SPDLOG_ERROR("Ignoring checkpoint directory {}: {}", dir, e.what());
    }
}

void FaabricMain::writeCheckpoint()
{
    const std::string& dir = faabric::util::getSystemConfig().checkpointDir;
    if (dir.empty()) {
        return;
    }
    try {
        size_t n = faabric::snapshot::getSnapshotRegistry().checkpointToDir(dir);
        SPDLOG_INFO("Checkpointed {} snapshots to {}", n, dir);
    } catch (const std::exception& e) {
        SPDLOG_ERROR("Checkpoint to {} failed: {}", dir, e.what());
    }
}

void FaabricMain::startRunner()
{
    faabric::planner::getPlannerClient().ping();
    // The planner answers: let it arbitrate state mains so that every worker
    // process agrees on them
    faabric::state::getInMemoryStateRegistry().setShared(true);
    faabric::scheduler::getScheduler().addHostToGlobalSet();
}

void FaabricMain::startStateServer()
{
    const std::string& mode = faabric::util::getSystemConfig().stateMode;
    if (mode != "inmemory") {
        // Values live in the external store: nothing to serve
        SPDLOG_INFO("Not starting state server in state mode {}", mode);
        return;
    }
    SPDLOG_INFO("Starting state server");
    stateServer.start();
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

void FaabricMain::startFunctionCallServer()
{
    SPDLOG_INFO("Starting function call server");
    functionServer.start();
}

void FaabricMain::shutdown()
{
    faabric::state::getInMemoryStateRegistry().setShared(false);
    SPDLOG_INFO("Leaving the planner's host set");
    faabric::scheduler::getScheduler().shutdown();

    // Stop taking work first, the services it depends on last
    SPDLOG_INFO("Stopping servers");
    functionServer.stop();
    pointToPointServer.stop();
    snapshotServer.stop();
    stateServer.stop();
    // nothing can change the images any more
    writeCheckpoint();
    SPDLOG_INFO("Worker shut down");
}

}
