#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/PointToPointCall.h>

#include <memory>
#include <string>
#include <vector>

namespace faabric::transport {

// Mock-mode capture (reference: src/transport/PointToPointClient.cpp:11-49)
std::vector<std::pair<std::string, faabric::PointToPointMappings>>
getSentMappings();

std::vector<std::pair<std::string, faabric::PointToPointMessage>>
getSentPointToPointMessages();

std::vector<std::tuple<std::string,
                       faabric::transport::PointToPointCall,
                       faabric::PointToPointMessage>>
getSentLockMessages();

void clearSentMessages();

class PointToPointClient : public faabric::transport::MessageEndpointClient
{
  public:
    explicit PointToPointClient(const std::string& hostIn);

    void sendMappings(faabric::PointToPointMappings& mappings);

    void sendMessage(const faabric::PointToPointMessage& msg,
                     int sequenceNum = NO_SEQUENCE_NUM);

    void groupLock(int appId, int groupId, int groupIdx, bool recursive = false);

    void groupUnlock(int appId,
                     int groupId,
                     int groupIdx,
                     bool recursive = false);

  private:
    void makeCoordinationRequest(int appId,
                                 int groupId,
                                 int groupIdx,
                                 faabric::transport::PointToPointCall call);
};

// Per-thread cached client for a host
std::shared_ptr<PointToPointClient> getPointToPointClient(
  const std::string& host);

void clearPointToPointClients();

}
