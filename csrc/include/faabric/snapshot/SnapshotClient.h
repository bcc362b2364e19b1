#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotApi.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/util/snapshot.h>

#include <memory>
#include <string>
#include <vector>

namespace faabric::snapshot {

// -----------------------------------
// Mocking (reference: src/snapshot/SnapshotClient.cpp:18-64)
// -----------------------------------
struct MockSnapshotUpdate
{
    std::vector<faabric::util::SnapshotDiff> diffs;
    // diffs above are non-owning: the payloads are kept here
    std::vector<std::vector<uint8_t>> diffData;
    std::vector<faabric::util::SnapshotMergeRegion> mergeRegions;
};

std::vector<
  std::pair<std::string, std::shared_ptr<faabric::util::SnapshotData>>>
getSnapshotPushes();

std::vector<std::pair<std::string, std::shared_ptr<MockSnapshotUpdate>>>
getSnapshotDiffPushes();

std::vector<std::pair<std::string, std::string>> getSnapshotDeletes();

std::vector<std::pair<std::string, std::tuple<int, int, std::string, int>>>
getThreadResults();

void clearMockSnapshotRequests();

// -----------------------------------
// Client
// -----------------------------------
class SnapshotClient final : public faabric::transport::MessageEndpointClient
{
  public:
    explicit SnapshotClient(const std::string& hostIn);

    void pushSnapshot(const std::string& key,
                      std::shared_ptr<faabric::util::SnapshotData> data);

    void pushSnapshotUpdate(
      std::string snapshotKey,
      const std::shared_ptr<faabric::util::SnapshotData>& data,
      const std::vector<faabric::util::SnapshotDiff>& diffs);

    void deleteSnapshot(const std::string& key);

    void pushThreadResult(uint32_t appId,
                          uint32_t messageId,
                          int returnValue,
                          const std::string& key,
                          const std::vector<faabric::util::SnapshotDiff>& diffs);
};

std::shared_ptr<SnapshotClient> getSnapshotClient(const std::string& host);

void clearSnapshotClients();

}
