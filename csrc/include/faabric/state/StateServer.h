#pragma once

#include <faabric/state/State.h>
#include <faabric/transport/MessageEndpointServer.h>

namespace faabric::state {

class StateServer final : public faabric::transport::MessageEndpointServer
{
  public:
    explicit StateServer(State& stateIn);

  private:
    State& state;

    void logOperation(const std::string& op);

    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    // Sync methods
    std::string recvSize(std::span<const uint8_t> buffer);

    std::string recvPull(std::span<const uint8_t> buffer);

    std::string recvPush(std::span<const uint8_t> buffer);

    std::string recvAppend(std::span<const uint8_t> buffer);

    std::string recvPullAppended(std::span<const uint8_t> buffer);

    std::string recvClearAppended(std::span<const uint8_t> buffer);

    std::string recvDelete(std::span<const uint8_t> buffer);
};

}
