#pragma once

#include <faabric/snapshot/SnapshotApi.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/MessageEndpointServer.h>

namespace faabric::snapshot {

class SnapshotServer final : public faabric::transport::MessageEndpointServer
{
  public:
    SnapshotServer();

  protected:
    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    std::string recvPushSnapshot(std::span<const uint8_t> buffer);

    std::string recvPushSnapshotUpdate(std::span<const uint8_t> buffer);

    std::string recvThreadResult(transport::Message& message);

    void recvDeleteSnapshot(std::span<const uint8_t> buffer);

  private:
    faabric::snapshot::SnapshotRegistry& reg;
};

}
