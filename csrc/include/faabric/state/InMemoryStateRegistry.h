#pragma once

#include <shared_mutex>
#include <string>
#include <unordered_map>

namespace faabric::state {

// Which host is the main (authoritative) copy of a key: the first to claim it.
// The reference keeps this in Redis under a lock
// (src/state/InMemoryStateRegistry.cpp:26-119); here it is the in-process
// Redis-compatible store, so the same protocol runs without a server.
class InMemoryStateRegistry
{
  public:
    InMemoryStateRegistry() = default;

    std::string getMasterIP(const std::string& user,
                            const std::string& key,
                            const std::string& thisIP,
                            bool claim);

    std::string getMasterIPForOtherMaster(const std::string& userIn,
                                          const std::string& keyIn,
                                          const std::string& thisIP);

    void clear();

  private:
    std::unordered_map<std::string, std::string> mainMap;
    std::shared_mutex mainMapMutex;
};

InMemoryStateRegistry& getInMemoryStateRegistry();

}
