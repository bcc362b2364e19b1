// Fault injection for the control plane (the reference has none).
// Rules match outgoing RPCs by destination port and/or message code and drop,
// delay or fail them.  Armed from code (tests) or from the environment:
//
//   FAABRIC_FAULTS="drop:port=8005,header=1,count=2;delay:port=8011,ms=50;error:header=7"
//
// The check on the hot path is one relaxed atomic load when nothing is armed.
#pragma once

#include <atomic>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

namespace faabric::util {

enum class FaultAction
{
    DROP,  // async: silently lost; sync: the caller sees a timeout
    DELAY, // sleep before sending
    ERROR, // the send throws
};

struct FaultRule
{
    FaultAction action = FaultAction::DROP;
    int port = -1;   // -1 = any
    int header = -1; // -1 = any
    int delayMs = 0;
    int count = -1; // how many times it fires (-1 = forever)
};

class FaultInjector
{
  public:
    static FaultInjector& get();

    void addRule(const FaultRule& rule);

    // "action:key=value,..;action:.." (see above)
    void addRulesFromString(const std::string& spec);

    void clear();

    bool armed() const { return nArmed.load(std::memory_order_relaxed) > 0; }

    // The rule that fires for this send, if any (consumes one of its counts)
    std::optional<FaultRule> match(int port, int header);

    long firedCount() const { return fired.load(); }

  private:
    FaultInjector();

    std::mutex mx;
    std::vector<FaultRule> rules;
    std::atomic<int> nArmed{ 0 };
    std::atomic<long> fired{ 0 };
};

}
