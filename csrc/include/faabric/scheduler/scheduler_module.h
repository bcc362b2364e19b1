// Per-host scheduler and the function-call RPCs.
//
// One header per module: the per-class headers of the reference's layout
// (faabric/scheduler/*.h) forward here, so either include style works.
#pragma once

#include <faabric/batch-scheduler/SchedulingDecision.h>
#include <faabric/planner/PlannerClient.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/Message.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/util/PeriodicBackgroundThread.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/snapshot.h>

#include <future>
#include <memory>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

// ==========================================================================
// scheduler/FunctionCallApi.h
// ==========================================================================
namespace faabric::scheduler {
enum FunctionCalls
{
    NoFunctionCall = 0,
    ExecuteFunctions = 1,
    Flush = 2,
    SetMessageResult = 3,
};
}

// ==========================================================================
// scheduler/FunctionCallClient.h
// ==========================================================================
namespace faabric::scheduler {

// -----------------------------------
// Mocking (reference: src/scheduler/FunctionCallClient.cpp:16-60)
// -----------------------------------
std::vector<std::pair<std::string, faabric::Message>> getFunctionCalls();

std::vector<std::pair<std::string, faabric::EmptyRequest>> getFlushCalls();

std::vector<
  std::pair<std::string, std::shared_ptr<faabric::BatchExecuteRequest>>>
getBatchRequests();

std::vector<std::pair<std::string, std::shared_ptr<faabric::Message>>>
getMessageResults();

void clearMockRequests();

// -----------------------------------
// Call client
// -----------------------------------
class FunctionCallClient : public faabric::transport::MessageEndpointClient
{
  public:
    explicit FunctionCallClient(const std::string& hostIn);

    void sendFlush();

    void executeFunctions(std::shared_ptr<faabric::BatchExecuteRequest> req);

    void setMessageResult(std::shared_ptr<faabric::Message> msg);
};

// -----------------------------------
// Static client pool
// -----------------------------------
std::shared_ptr<FunctionCallClient> getFunctionCallClient(
  const std::string& otherHost);

void clearFunctionCallClients();

}

// ==========================================================================
// scheduler/FunctionCallServer.h
// ==========================================================================
namespace faabric::scheduler {

class Scheduler;

class FunctionCallServer final
  : public faabric::transport::MessageEndpointServer
{
  public:
    FunctionCallServer();

  private:
    Scheduler& scheduler;

    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    std::string recvFlush(std::span<const uint8_t> buffer);

    void recvExecuteFunctions(std::span<const uint8_t> buffer);

    void recvSetMessageResult(std::span<const uint8_t> buffer);
};

}

// ==========================================================================
// scheduler/Scheduler.h
// ==========================================================================
// Per-worker scheduler: owns the executors of this host (GPU worker), runs the
// batches the planner dispatches, tracks thread results and asks the planner
// for migration opportunities.  Reference: include/faabric/scheduler/
// Scheduler.h:34-140, src/scheduler/Scheduler.cpp:25-524.



#define AVAILABLE_HOST_SET "available_hosts"

namespace faabric::executor {
class Executor;
}

namespace faabric::scheduler {

class Scheduler;

Scheduler& getScheduler();

// Reaps executors that have been idle for longer than BOUND_TIMEOUT
class SchedulerReaperThread : public faabric::util::PeriodicBackgroundThread
{
  public:
    void doWork() override;
};

class Scheduler
{
  public:
    Scheduler();

    ~Scheduler();

    // Key of the warm-executor pools (function, plus the virtual host)
    static std::string executorKeyFor(const faabric::Message& msg);

    void executeBatch(std::shared_ptr<faabric::BatchExecuteRequest> req);

    void reset();

    void resetThreadLocalCache();

    void shutdown();

    bool isShutdown() { return _isShutdown; }

    long getFunctionExecutorCount(const faabric::Message& msg);

    void flushLocally();

    // ----------------------------------
    // Message results (threads)
    // ----------------------------------
    void setThreadResultLocally(uint32_t appId,
                                uint32_t msgId,
                                int32_t returnValue,
                                faabric::transport::Message& message);

    // The last argument keeps diff payloads alive until the result is consumed
    std::vector<std::pair<uint32_t, int32_t>> awaitThreadResults(
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      int timeoutMs = DEFAULT_THREAD_RESULT_TIMEOUT_MS);

    size_t getCachedMessageCount();

    std::string getThisHost();

    void addHostToGlobalSet();

    void addHostToGlobalSet(
      const std::string& host,
      std::shared_ptr<faabric::HostResources> overwriteResources = nullptr);

    void removeHostFromGlobalSet(const std::string& host);

    void setThisHostResources(faabric::HostResources& res);

    // ----------------------------------
    // Testing
    // ----------------------------------
    std::vector<faabric::Message> getRecordedMessages();

    void clearRecordedMessages();

    // ----------------------------------
    // Function Migration
    // ----------------------------------
    std::shared_ptr<faabric::PendingMigration> checkForMigrationOpportunities(
      faabric::Message& msg,
      int overwriteNewGroupId = 0);

    // Idle-executor reaping, returns how many were reaped
    // Deletes a snapshot on every other host registered with the planner
    // (declared by the reference, include/faabric/scheduler/Scheduler.h:51)
    void broadcastSnapshotDelete(const faabric::Message& msg, const std::string& snapshotKey);

    int reapStaleExecutors();

    // Called by an executor when it becomes claimable again
    void notifyExecutorIdle(const std::string& funcKey, std::weak_ptr<faabric::executor::Executor> executor);

    static const int DEFAULT_THREAD_RESULT_TIMEOUT_MS = 20000;

  private:
    std::string thisHost;

    faabric::util::SystemConfig& conf;

    std::shared_mutex mx;

    std::atomic<bool> _isShutdown = false;

    // ---- Executors ----
    std::unordered_map<std::string,
                       std::vector<std::shared_ptr<faabric::executor::Executor>>>
      executors;

    // ---- Threads ----
    // Recently released executors per function (hints: entries may be stale)
    // A sleeping mutex by default.  FAABRIC_SCHED_IDLE_LOCK=spin makes waiters
    // spin instead: measured on a 128-core box with 1024 pool threads that is
    // SLOWER (fan-out 6.8 ms vs 3.3 ms) - the releasers of one host's batch
    // starve the claims of the next host's
    struct IdleLock
    {
        std::atomic_flag flag = ATOMIC_FLAG_INIT;
        std::mutex mx;
        bool spin = false;

        IdleLock()
        {
            const char* v = getenv("FAABRIC_SCHED_IDLE_LOCK");
            spin = v != nullptr && std::string(v) == "spin";
        }

        void lock()
        {
            if (!spin) {
                mx.lock();
                return;
            }
            for (int spins = 0; flag.test_and_set(std::memory_order_acquire); spins++) {
                if (spins < 64) {
                    __builtin_ia32_pause();
                } else {
                    std::this_thread::yield();
                }
            }
        }

        void unlock()
        {
            if (!spin) {
                mx.unlock();
                return;
            }
            flag.clear(std::memory_order_release);
        }
    };
    IdleLock idleMx;
    std::unordered_map<std::string, std::vector<std::weak_ptr<faabric::executor::Executor>>> idleExecutors;

    faabric::snapshot::SnapshotRegistry& reg;

    std::unordered_map<uint32_t, faabric::transport::Message>
      threadResultMessages;
    std::mutex threadResultsMx;

    // ---- Planner ----
    faabric::planner::KeepAliveThread keepAliveThread;
    bool keepAliveRunning = false;
    std::set<std::string> servedHosts;

    // ---- Actual scheduling ----
    SchedulerReaperThread reaperThread;

    std::shared_ptr<faabric::executor::Executor> claimExecutor(
      faabric::Message& msg,
      std::unique_lock<std::shared_mutex>& schedulerLock);

    std::shared_ptr<faabric::executor::Executor> claimExecutorForKey(
      const std::string& key,
      faabric::Message& msg,
      std::unique_lock<std::shared_mutex>& schedulerLock);

    // ---- Point-to-point ----
    faabric::transport::PointToPointBroker& broker;

    // ---- Mock records ----
    std::vector<faabric::Message> recordedMessages;
};

}

