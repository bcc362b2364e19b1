#pragma once

#include <faabric/executor/ExecutorFactory.h>
#include <faabric/scheduler/FunctionCallServer.h>
#include <faabric/scheduler/Scheduler.h>
#include <faabric/snapshot/SnapshotServer.h>
#include <faabric/state/StateServer.h>
#include <faabric/transport/PointToPointServer.h>
#include <faabric/util/config.h>

namespace faabric::runner {

// Boots a worker: crash handler, registers the host with the planner, starts
// the State / Snapshot / PointToPoint / FunctionCall servers (reference:
// src/runner/FaabricMain.cpp:11-109).  On a GPU box it also initialises CUDA
// and binds the worker to its GPU.
class FaabricMain
{
  public:
    explicit FaabricMain(std::shared_ptr<faabric::executor::ExecutorFactory> fac);

    void startBackground();

    void startRunner();

    void startFunctionCallServer();

    void startStateServer();

    void startSnapshotServer();

    void startPointToPointServer();

    void shutdown();

  private:
    faabric::state::StateServer stateServer;
    faabric::scheduler::FunctionCallServer functionServer;
    faabric::snapshot::SnapshotServer snapshotServer;
    faabric::transport::PointToPointServer pointToPointServer;
};

}
