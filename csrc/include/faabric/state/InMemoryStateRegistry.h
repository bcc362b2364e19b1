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

    // Shared mode: mains are elected by the planner, so every worker process
    // agrees (set by FaabricMain once the planner answers).  Otherwise the
    // in-process key-value emulation arbitrates (single process, tests).
    void setShared(bool shared) { sharedViaPlanner = shared; }

    bool isShared() const { return sharedViaPlanner; }

    // Forgets who the main of user/key is: locally only, or in the shared
    // store as well (the main itself deleting the value)
    void dropMain(const std::string& user, const std::string& key, bool everywhere);

  private:
    std::unordered_map<std::string, std::string> mainMap;
    std::shared_mutex mainMapMutex;
    bool sharedViaPlanner = false;
};

InMemoryStateRegistry& getInMemoryStateRegistry();

}
