#pragma once

#include <faabric/proto/faabric.pb.h>

#include <memory>
#include <stdexcept>

namespace faabric::executor {

class Executor;

class ExecutorContextException : public std::runtime_error
{
  public:
    explicit ExecutorContextException(const std::string& message)
      : std::runtime_error(message)
    {}
};

// Thread-local handle on "what am I executing": set around executeTask so
// library code (MPI shim, chaining, state) can find the current message
class ExecutorContext
{
  public:
    ExecutorContext(Executor* executorIn,
                    std::shared_ptr<faabric::BatchExecuteRequest> reqIn,
                    int msgIdx);

    static bool isSet();

    static void set(Executor* executorIn,
                    std::shared_ptr<faabric::BatchExecuteRequest> reqIn,
                    int msgIdxIn);

    static void unset();

    static std::shared_ptr<ExecutorContext> get();

    Executor* getExecutor() { return executor; }

    std::shared_ptr<faabric::BatchExecuteRequest> getBatchRequest()
    {
        return req;
    }

    faabric::Message& getMsg()
    {
        if (req == nullptr) {
            throw ExecutorContextException("Getting message when no request set in context");
        }
        return *req->mutable_messages(msgIdx);
    }

    int getMsgIdx() const { return msgIdx; }

  private:
    Executor* executor = nullptr;
    std::shared_ptr<faabric::BatchExecuteRequest> req = nullptr;
    int msgIdx = 0;
};

}
