#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/scheduler/FunctionCallApi.h>
#include <faabric/transport/MessageEndpointClient.h>

#include <memory>
#include <string>
#include <vector>

namespace faabric::scheduler {

// -----------------------------------
// Mocking (reference: src/scheduler/FunctionCallClient.cpp:16-60)
// -----------------------------------
std::vector<std::pair<std::string, faabric::Message>> getFunctionCalls();

std::vector<std::pair<std::string, faabric::EmptyRequest>> getFlushCalls();

std::vector<
  std::pair<std::string, std::shared_ptr<faabric::BatchExecuteRequest>>>
getBatchRequests();

std::vector<std::pair<std::string, std::shared_ptr<faabric::Message>>>
getMessageResults();

void clearMockRequests();

// -----------------------------------
// Call client
// -----------------------------------
class FunctionCallClient : public faabric::transport::MessageEndpointClient
{
  public:
    explicit FunctionCallClient(const std::string& hostIn);

    void sendFlush();

    void executeFunctions(std::shared_ptr<faabric::BatchExecuteRequest> req);

    void setMessageResult(std::shared_ptr<faabric::Message> msg);
};

// -----------------------------------
// Static client pool
// -----------------------------------
std::shared_ptr<FunctionCallClient> getFunctionCallClient(
  const std::string& otherHost);

void clearFunctionCallClients();

}
