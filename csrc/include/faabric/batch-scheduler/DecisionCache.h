// Remembers where an app of a given size was placed so a repeat invocation can
// skip scheduling (reference: src/batch-scheduler/DecisionCache.cpp:7-78)
#pragma once

#include <faabric/batch-scheduler/SchedulingDecision.h>

#include <memory>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace faabric::batch_scheduler {

class CachedDecision
{
  public:
    CachedDecision(const std::vector<std::string>& hostsIn, int groupIdIn);

    std::vector<std::string> getHosts() { return hosts; }

    int getGroupId() const { return groupId; }

  private:
    std::vector<std::string> hosts;
    int groupId = 0;
};

class DecisionCache
{
  public:
    std::shared_ptr<CachedDecision> getCachedDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    void addCachedDecision(std::shared_ptr<faabric::BatchExecuteRequest> req,
                           SchedulingDecision& decision);

    void clear();

  private:
    std::string getCacheKey(std::shared_ptr<faabric::BatchExecuteRequest> req);

    std::shared_mutex mx;
    std::unordered_map<std::string, std::shared_ptr<CachedDecision>>
      cachedDecisions;
};

DecisionCache& getSchedulingDecisionCache();

}
