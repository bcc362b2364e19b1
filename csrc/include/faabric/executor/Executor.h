// Executor: a warm container for one (user, function, app) that runs the
// messages of a batch on a pool of threads.  Users subclass it and implement
// executeTask (reference: include/faabric/executor/Executor.h:21-118,
// src/executor/Executor.cpp:38-743).  Additions for GPUs: every executor is
// bound to a GPU (device id + compute stream) and may expose a device memory
// view that is snapshot / restored / diff-pushed with the device kernels.
#pragma once

#include <faabric/executor/ExecutorTask.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/util/clock.h>
#include <faabric/util/dirty.h>
#include <faabric/util/exception.h>
#include <faabric/util/hwloc.h>
#include <faabric/util/queue.h>
#include <faabric/util/snapshot.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <span>
#include <thread>
#include <vector>

#define POOL_SHUTDOWN -1

namespace faabric::executor {

class ChainedCallException : public faabric::util::FaabricException
{
  public:
    explicit ChainedCallException(std::string message)
      : FaabricException(std::move(message))
    {}
};

class Executor : public std::enable_shared_from_this<Executor>
{
  public:
    std::string id;

    explicit Executor(faabric::Message& msg);

    virtual ~Executor();

    // Must be called before the executor is destroyed
    void shutdown();

    std::vector<std::pair<uint32_t, int32_t>> executeThreads(
      std::shared_ptr<faabric::BatchExecuteRequest> req,
      const std::vector<faabric::util::SnapshotMergeRegion>& mergeRegions);

    void executeTasks(std::vector<int> msgIdxs,
                      std::shared_ptr<faabric::BatchExecuteRequest> req);

    // ---- hooks for subclasses ----
    virtual void reset(faabric::Message& msg);

    virtual int32_t executeTask(
      int threadPoolIdx,
      int msgIdx,
      std::shared_ptr<faabric::BatchExecuteRequest> req);

    virtual std::span<uint8_t> getMemoryView();

    virtual void restore(const std::string& snapshotKey);

    virtual void setMemorySize(size_t newSize);

    virtual size_t getMaxMemorySize();

    // ---- GPU binding ----
    // App this executor is currently working for (0 when idle)
    int getCurrentAppId() const { return currentAppId.load(); }

    int getGpuIdx() const { return gpuIdx; }

    // Compute stream of this executor (cudaStream_t); nullptr without a GPU
    void* getComputeStream() const { return computeStream; }

    // ---- claiming ----
    bool tryClaim();

    void claim();

    void releaseClaim();

    bool isExecuting();

    bool isShutdown() { return _isShutdown; }

    long getMillisSinceLastExec();

    std::shared_ptr<faabric::util::SnapshotData> getMainThreadSnapshot(
      faabric::Message& msg,
      bool createIfNotExists = false);

    // ---- chained calls ----
    void addChainedMessage(const faabric::Message& msg);

    const faabric::Message& getChainedMessage(int messageId);

    std::set<unsigned int> getChainedMessageIds();

    std::vector<faabric::util::SnapshotDiff> mergeDirtyRegions(
      const faabric::Message& msg,
      const std::vector<char>& extraDirtyPages = {});

    // Blocks until every pool thread finished (tests)
    void joinThreadPool();

  protected:
    virtual void setUpThreadPool();

    faabric::Message boundMessage;

    faabric::snapshot::SnapshotRegistry& reg;

    std::shared_ptr<faabric::util::DirtyTracker> tracker;

    uint32_t threadPoolSize = 0;

    std::map<int, std::shared_ptr<faabric::BatchExecuteRequest>> chainedMessages;

  private:
    std::atomic<bool> claimed = false;

    std::atomic<bool> _isShutdown = false;

    std::atomic<int> batchCounter = 0;
    std::atomic<int> currentAppId = 0;

    std::atomic<int> threadBatchCounter = 0;

    faabric::util::TimePoint lastExec;

    // ---- Application threads ----
    std::shared_mutex threadExecutionMutex;
    std::vector<char> dirtyRegions;
    std::vector<std::vector<char>> threadLocalDirtyRegions;
    void deleteMainThreadSnapshot(const faabric::Message& msg);

    // ---- Function execution thread pool ----
    std::mutex threadsMutex;
    std::vector<std::shared_ptr<std::jthread>> threadPoolThreads;
    std::set<int> availablePoolThreads;

    std::vector<faabric::util::Queue<ExecutorTask>> threadTaskQueues;

    std::mutex chainedMx;

    int gpuIdx = -1;
    void* computeStream = nullptr;

    void threadPoolThread(std::stop_token st, int threadPoolIdx);
};

}
