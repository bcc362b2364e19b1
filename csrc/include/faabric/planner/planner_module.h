// Planner: state, RPC codes, client, server, HTTP handler.
//
// One header per module: the per-class headers of the reference's layout
// (faabric/planner/*.h) forward here, so either include style works.
#pragma once

#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/util/PeriodicBackgroundThread.h>

#include <atomic>
#include <condition_variable>
#include <future>
#include <map>
#include <memory>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_set>
#include <vector>
#include <functional>
#include <thread>

// ==========================================================================
// (merged) faabric/endpoint
// ==========================================================================
namespace faabric::endpoint {

struct HttpRequest
{
    std::string method;
    std::string target;
    std::map<std::string, std::string> headers;
    std::string body;
};

struct HttpResponse
{
    int status = 200;
    std::string body;
    std::map<std::string, std::string> headers;
};

class HttpRequestHandler
{
  public:
    virtual ~HttpRequestHandler() = default;

    virtual void onRequest(const HttpRequest& request, HttpResponse& response) = 0;
};

enum class EndpointMode
{
    SIGNAL,
    BG_THREAD
};

class FaabricEndpoint
{
  public:
    FaabricEndpoint();

    FaabricEndpoint(int port,
                    int threadCount,
                    std::shared_ptr<HttpRequestHandler> requestHandlerIn);

    FaabricEndpoint(const FaabricEndpoint&) = delete;

    ~FaabricEndpoint();

    // SIGNAL: blocks until SIGINT/SIGTERM; BG_THREAD: returns immediately
    void start(EndpointMode mode = EndpointMode::SIGNAL);

    void stop();

    int getPort() const { return port.load(); }

  private:
    // (read by other threads while start() is binding an ephemeral port)
    std::atomic<int> port;
    int threadCount;
    std::shared_ptr<HttpRequestHandler> requestHandler;
    struct Impl;
    std::unique_ptr<Impl> impl;
};

}

// ==========================================================================
// endpoint/FaabricEndpointHandler.h
// ==========================================================================
namespace faabric::endpoint {

// Worker-side handler: workers do not take HTTP requests (the planner does),
// everything is rejected (reference: src/endpoint/FaabricEndpointHandler.cpp)
class FaabricEndpointHandler final : public HttpRequestHandler
{
  public:
    void onRequest(const HttpRequest& request, HttpResponse& response) override;
};

}



// ==========================================================================
// planner/PlannerState.h
// ==========================================================================
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
    // appId -> (message id -> position in the app's scheduling decision)
    std::map<int, std::unordered_map<int, int>> inFlightIdPos;

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

// ==========================================================================
// planner/Planner.h
// ==========================================================================
// The planner: host membership, batch scheduling, result storage, migration
// and freeze/thaw (reference: include/faabric/planner/Planner.h:23-145,
// src/planner/Planner.cpp).  Hosts are GPU workers of the box.



namespace faabric::planner {

enum FlushType
{
    NoFlushType = 0,
    Hosts = 1,
    Executors = 2,
    SchedulingState = 3,
};

class Planner
{
  public:
    Planner();

    // ----------
    // Planner config
    // ----------
    PlannerConfig getConfig();

    // Seconds without a keep-alive after which a host is dropped
    void setHostKeepAliveTimeout(int seconds);

    void printConfig() const;

    std::string getPolicy();

    void setPolicy(const std::string& newPolicy);

    // ----------
    // Util public API
    // ----------
    bool reset();

    bool flush(faabric::planner::FlushType flushType);

    // ----------
    // Host membership public API
    // ----------
    std::vector<std::shared_ptr<Host>> getAvailableHosts();

    bool registerHost(const Host& hostIn, bool overwrite);

    // Best effort
    void removeHost(const Host& hostIn);

    // ----------
    // Request scheduling public API
    // ----------
    void setMessageResult(std::shared_ptr<faabric::Message> msg);

    // Many results under one acquisition of the planner's lock
    void setMessageResults(const std::vector<std::shared_ptr<faabric::Message>>& msgs);

    // Combining entry point for result producers in the planner's process:
    // the result is recorded by this call or by a concurrent caller that is
    // already draining (returns at once in that case)
    void submitMessageResult(std::shared_ptr<faabric::Message> msg);

    // Non-blocking: nullptr if not ready (and the caller is registered as a
    // waiter when it named its main host)
    std::shared_ptr<faabric::Message> getMessageResult(
      std::shared_ptr<faabric::Message> msg);

    void preloadSchedulingDecision(
      int appId,
      std::shared_ptr<batch_scheduler::SchedulingDecision> decision);

    std::shared_ptr<faabric::BatchExecuteRequestStatus> getBatchResults(
      int32_t appId);

    std::shared_ptr<faabric::batch_scheduler::SchedulingDecision>
    getSchedulingDecision(std::shared_ptr<BatchExecuteRequest> req);

    faabric::batch_scheduler::InFlightReqs getInFlightReqs();

    // Blocks until the app has no message in flight (false on timeout).
    // For callers living in the planner's process.
    bool waitForAppToFinish(int32_t appId, int timeoutMs);

    int getNumMigrations();

    std::set<std::string> getNextEvictedHostIps();

    std::map<int32_t, std::shared_ptr<BatchExecuteRequest>> getEvictedReqs();

    // The main entry point: schedule + dispatch
    std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> callBatch(
      std::shared_ptr<BatchExecuteRequest> req);

    // Spot policy: which hosts go away next
    void setNextEvictedVm(const std::set<std::string>& vmIps);

    // State main election: returns the main host of user/key, electing `host`
    // if there is none and `claim` is set ("" = none); `drop` forgets it
    std::string stateMain(const std::string& user, const std::string& key, const std::string& host, bool claim, bool drop);

  private:
    std::shared_mutex plannerMx;

    std::mutex pendingResultsMx;
    std::vector<std::shared_ptr<faabric::Message>> pendingResults;
    bool drainingResults = false;

    // Look-ups shared by consecutive results of one app / one host
    struct ResultContext
    {
        bool valid = false;
        int appId = 0;
        std::map<int, std::shared_ptr<faabric::Message>>* results = nullptr;
        std::string hostName;
        std::shared_ptr<Host> host;
        bool hostKnown = false;
    };

    // Caller holds plannerMx exclusively
    void recordResultLocked(const std::shared_ptr<faabric::Message>& msg,
                            std::vector<std::string>& toNotify,
                            ResultContext* ctx = nullptr);
    std::condition_variable_any appFinishedCv;

    void compactInFlightLocked();

    PlannerState state;
    PlannerConfig config;

    faabric::snapshot::SnapshotRegistry& snapshotRegistry;

    // ----------
    // Util private API
    // ----------
    void flushHosts();

    void flushExecutors();

    void flushSchedulingState();

    // ----------
    // Host membership private API
    // ----------
    bool isHostExpired(std::shared_ptr<Host> host, long epochTimeMs = 0);

    // ----------
    // Request scheduling private API
    // ----------
    std::shared_ptr<batch_scheduler::SchedulingDecision>
    getPreloadedSchedulingDecision(
      int32_t appId,
      std::shared_ptr<BatchExecuteRequest> ber);

    void dispatchSchedulingDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> decision);
};

Planner& getPlanner();

}

// ==========================================================================
// planner/PlannerApi.h
// ==========================================================================
namespace faabric::planner {
enum PlannerCalls
{
    NoPlanerCall = 0,
    // Util
    Ping = 1,
    // Host-membership calls
    GetAvailableHosts = 2,
    RegisterHost = 3,
    RemoveHost = 4,
    // Scheduling calls
    SetMessageResult = 8,
    GetMessageResult = 9,
    GetBatchResults = 10,
    GetSchedulingDecision = 11,
    GetNumMigrations = 12,
    CallBatch = 13,
    PreloadSchedulingDecision = 14,
    // Shared registry of state mains (replaces the reference's Redis keys)
    StateMain = 20,
};
}

// ==========================================================================
// planner/PlannerClient.h
// ==========================================================================
namespace faabric::planner {

// Re-registers this host with the planner every timeout/2
class KeepAliveThread : public faabric::util::PeriodicBackgroundThread
{
  public:
    void doWork() override;

    // Adds (or replaces) the keep-alive of one host served by this process:
    // this host itself and every per-GPU virtual host it exposes
    void setRequest(std::shared_ptr<RegisterHostRequest> hostReqIn);

    // Returns how many hosts are still being kept alive
    size_t removeRequest(const std::string& hostIp);

    std::shared_mutex keepAliveThreadMx;

  private:
    std::map<std::string, std::shared_ptr<RegisterHostRequest>> hostReqs;
};

struct PlannerCache
{
    std::unordered_map<uint32_t, std::promise<std::shared_ptr<faabric::Message>>>
      plannerResults;

    // Snapshots already pushed to the planner, by key
    std::set<std::string> pushedSnapshots;
};

class PlannerClient final : public faabric::transport::MessageEndpointClient
{
  public:
    PlannerClient();

    explicit PlannerClient(const std::string& plannerIp);

    // ------
    // Util
    // ------
    void ping();

    void clearCache();

    // ------
    // Host membership calls
    // ------
    std::vector<Host> getAvailableHosts();

    // Returns the keep-alive timeout (seconds)
    int registerHost(std::shared_ptr<RegisterHostRequest> req);

    void removeHost(std::shared_ptr<RemoveHostRequest> req);

    // ------
    // Scheduling calls
    // ------
    void setMessageResult(std::shared_ptr<faabric::Message> msg);

    // Called by the FunctionCallServer when the planner notifies a result
    // onlyIfAwaited: deliver only to a wait that is already registered (the
    // planner's notifications, which are always answers to one)
    void setMessageResultLocally(std::shared_ptr<faabric::Message> msg, bool onlyIfAwaited = false);

    faabric::Message getMessageResult(int appId, int msgId, int timeoutMs);

    faabric::Message getMessageResult(const faabric::Message& msg,
                                      int timeoutMs);

    std::shared_ptr<faabric::BatchExecuteRequestStatus> getBatchResults(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    faabric::batch_scheduler::SchedulingDecision callFunctions(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    faabric::batch_scheduler::SchedulingDecision getSchedulingDecision(
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    int getNumMigrations();

    std::string stateMain(const std::string& user, const std::string& key, const std::string& host, bool claim, bool drop = false);

    void preloadSchedulingDecision(
      std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> preloadDec);

  private:
    std::mutex plannerCacheMx;
    PlannerCache cache;

    faabric::snapshot::SnapshotRegistry& snapshotRegistry;

    faabric::Message doGetMessageResult(
      std::shared_ptr<faabric::Message> msgPtr,
      int timeoutMs);
};

PlannerClient& getPlannerClient();

}

// ==========================================================================
// planner/PlannerEndpointHandler.h
// ==========================================================================
namespace faabric::planner {

// JSON-over-HTTP control API of the planner (reference:
// src/planner/PlannerEndpointHandler.cpp:15-421): body = HttpMessage JSON
class PlannerEndpointHandler final : public faabric::endpoint::HttpRequestHandler
{
  public:
    void onRequest(const faabric::endpoint::HttpRequest& request,
                   faabric::endpoint::HttpResponse& response) override;
};

}

// ==========================================================================
// planner/PlannerServer.h
// ==========================================================================
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

