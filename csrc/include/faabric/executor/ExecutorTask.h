#pragma once

#include <faabric/proto/faabric.pb.h>

#include <memory>

namespace faabric::executor {

class ExecutorTask
{
  public:
    ExecutorTask() = default;

    ExecutorTask(int messageIndexIn,
                 std::shared_ptr<faabric::BatchExecuteRequest> reqIn)
      : req(std::move(reqIn))
      , messageIndex(messageIndexIn)
    {}

    // Shutdown sentinel for pool threads
    static const int POOL_SHUTDOWN = -1;

    std::shared_ptr<faabric::BatchExecuteRequest> req;
    int messageIndex = 0;
};

}
