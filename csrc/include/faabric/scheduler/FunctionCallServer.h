#pragma once

#include <faabric/scheduler/FunctionCallApi.h>
#include <faabric/transport/MessageEndpointServer.h>

namespace faabric::scheduler {

class Scheduler;

class FunctionCallServer final
  : public faabric::transport::MessageEndpointServer
{
  public:
    FunctionCallServer();

  private:
    Scheduler& scheduler;

    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    std::string recvFlush(std::span<const uint8_t> buffer);

    void recvExecuteFunctions(std::span<const uint8_t> buffer);

    void recvSetMessageResult(std::span<const uint8_t> buffer);
};

}
