#pragma once

#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerApi.h>
#include <faabric/transport/MessageEndpointServer.h>

namespace faabric::planner {

class PlannerServer final : public faabric::transport::MessageEndpointServer
{
  public:
    PlannerServer();

  protected:
    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    // Asynchronous calls
    void recvSetMessageResult(std::span<const uint8_t> buffer);

    // Synchronous calls
    std::string recvPing();

    std::string recvGetAvailableHosts();

    std::string recvRegisterHost(std::span<const uint8_t> buffer);

    std::string recvRemoveHost(std::span<const uint8_t> buffer);

    std::string recvGetMessageResult(std::span<const uint8_t> buffer);

    std::string recvGetBatchResults(std::span<const uint8_t> buffer);

    std::string recvGetSchedulingDecision(std::span<const uint8_t> buffer);

    std::string recvGetNumMigrations(std::span<const uint8_t> buffer);

    std::string recvPreloadSchedulingDecision(std::span<const uint8_t> buffer);

    std::string recvCallBatch(std::span<const uint8_t> buffer);

    std::string recvStateMain(std::span<const uint8_t> buffer);

  private:
    faabric::planner::Planner& planner;
};

}
