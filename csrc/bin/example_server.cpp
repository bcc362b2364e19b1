// Minimal worker embedding the runtime: registers an executor factory whose
// functions just log and echo, then serves until signalled
// (reference: examples/server.cpp)
#include <faabric/endpoint/FaabricEndpoint.h>
#include <faabric/executor/ExecutorFactory.h>
#include <faabric/runner/FaabricMain.h>
#include <faabric/util/logging.h>

using namespace faabric::executor;

class ExampleExecutor : public Executor
{
  public:
    explicit ExampleExecutor(faabric::Message& msg)
      : Executor(msg)
    {}

    int32_t executeTask(int threadPoolIdx, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req) override
    {
        auto& msg = *req->mutable_messages(msgIdx);
        SPDLOG_INFO("Hello world! (executing {}/{} id {})", msg.user(), msg.function(), msg.id());
        msg.set_outputdata("This is hello output!");
        return 0;
    }
};

class ExampleExecutorFactory : public ExecutorFactory
{
  protected:
    std::shared_ptr<Executor> createExecutor(faabric::Message& msg) override
    {
        return std::make_shared<ExampleExecutor>(msg);
    }
};

int main()
{
    faabric::util::initLogging();
    SPDLOG_INFO("Starting faabric example worker");
    faabric::runner::FaabricMain m(std::make_shared<ExampleExecutorFactory>());
    m.startBackground();

    // Workers do not serve HTTP, but the endpoint gives us signal handling
    faabric::endpoint::FaabricEndpoint endpoint;
    endpoint.start(faabric::endpoint::EndpointMode::SIGNAL);

    SPDLOG_INFO("Shutting down example worker");
    m.shutdown();
    return 0;
}
