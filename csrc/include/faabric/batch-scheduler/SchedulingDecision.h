// Placement of the messages of one app onto hosts (GPUs).  Parallel vectors,
// one entry per message (reference: SchedulingDecision.h:58-119).
#pragma once

#include <faabric/proto/faabric.pb.h>

#include <cstdint>
#include <set>
#include <string>
#include <vector>

namespace faabric::batch_scheduler {

class SchedulingDecision
{
  public:
    static SchedulingDecision fromPointToPointMappings(
      faabric::PointToPointMappings& mappings);

    SchedulingDecision(uint32_t appIdIn, int32_t groupIdIn);

    bool operator==(const SchedulingDecision& rhs) const = default;

    uint32_t appId = 0;

    int32_t groupId = 0;

    int32_t nFunctions = 0;

    std::vector<std::string> hosts;

    std::vector<int32_t> messageIds;

    std::vector<int32_t> appIdxs;

    std::vector<int32_t> groupIdxs;

    // "MPI port" of each message: on GPUs the mailbox / stream slot index
    std::vector<int32_t> mpiPorts;

    std::string returnHost;

    // True when every message sits on one host AND that host is this one
    bool isSingleHost() const;

    void addMessage(const std::string& host, const faabric::Message& msg);

    void addMessage(const std::string& host, int32_t messageId, int32_t appIdx);

    void addMessage(const std::string& host,
                    int32_t messageId,
                    int32_t appIdx,
                    int32_t groupIdx);

    void addMessageInPosition(int32_t pos,
                              const std::string& host,
                              int32_t messageId,
                              int32_t appIdx,
                              int32_t groupIdx,
                              int32_t mpiPort);

    // Returns the MPI port the message vacated
    int32_t removeMessage(int32_t messageId);

    std::set<std::string> uniqueHosts();

    void print(const std::string& logLevel = "debug");

    std::string toString() const;
};

}
