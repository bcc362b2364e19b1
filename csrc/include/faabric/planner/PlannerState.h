#pragma once

#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/proto/faabric.pb.h>

#include <atomic>
#include <map>
#include <memory>
#include <set>
#include <unordered_set>
#include <string>
#include <vector>

namespace faabric::planner {

// Everything the planner knows (reference: include/faabric/planner/
// PlannerState.h:13-57)
struct PlannerState
{
    // Scheduling policy (bin-pack | compact | spot)
    std::string policy;

    // Registered hosts (= GPU workers), by address
    std::map<std::string, std::shared_ptr<Host>> hostMap;

    // appId -> msgId -> finished message
    std::map<int, std::map<int, std::shared_ptr<faabric::Message>>> appResults;

    // msgId -> hosts waiting to be told about its result
    std::map<int, std::vector<std::string>> appResultWaiters;

    // In-flight apps: request (messages still running) + current placement
    faabric::batch_scheduler::InFlightReqs inFlightReqs;

    // Messages that have finished but are still physically present in
    // inFlightReqs: results are recorded in O(1) and the request/decision
    // vectors are compacted in one pass before anybody reads them
    std::map<int, std::unordered_set<int>> finishedInFlight;

    // Placements fixed ahead of time (MPI / OpenMP two-step creation, tests)
    std::map<int, std::shared_ptr<batch_scheduler::SchedulingDecision>>
      preloadedSchedulingDecisions;

    std::atomic<int> numMigrations = 0;

    // Apps frozen by a spot eviction, waiting for capacity
    std::map<int, std::shared_ptr<BatchExecuteRequest>> evictedRequests;

    // Main host of every in-memory state value (user_key -> host)
    std::map<std::string, std::string> stateMains;

    // Hosts that will be evicted next (spot policy)
    std::set<std::string> nextEvictedHostIps;
};

}
