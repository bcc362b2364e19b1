// Distributed state: values, registry, client and server.
//
// One header per module: the per-class headers of the reference's layout
// (faabric/state/*.h) forward here, so either include style works.
#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/util/exception.h>
#include <faabric/util/memory.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <mutex>
#include <set>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <condition_variable>
#include <deque>
#include <map>
#include <unordered_set>

// ==========================================================================
// (merged) faabric/redis
// ==========================================================================
namespace faabric::redis {

enum RedisRole
{
    QUEUE,
    STATE,
};

class RedisNoResponseException : public std::runtime_error
{
  public:
    RedisNoResponseException()
      : std::runtime_error("No response from redis (timeout)")
    {}
};

class Redis
{
  public:
    ~Redis() = default;

    static Redis& getQueue();

    static Redis& getState();

    // ---- Standard Redis commands ----
    void ping();

    std::vector<uint8_t> get(const std::string& key);

    size_t strlen(const std::string& key);

    void get(const std::string& key, uint8_t* buffer, size_t size);

    void set(const std::string& key, const std::vector<uint8_t>& value);

    void set(const std::string& key, const uint8_t* value, size_t size);

    void del(const std::string& key);

    long getCounter(const std::string& key);

    long incr(const std::string& key);

    long decr(const std::string& key);

    long incrByLong(const std::string& key, long val);

    long decrByLong(const std::string& key, long val);

    void setRange(const std::string& key,
                  long offset,
                  const uint8_t* value,
                  size_t size);

    // Pipelined variant: queued then flushed (immediate here)
    void setRangePipeline(const std::string& key,
                          long offset,
                          const uint8_t* value,
                          size_t size);

    void flushPipeline(long pipelineLength);

    void getRange(const std::string& key,
                  uint8_t* buffer,
                  size_t bufferLen,
                  long start,
                  long end);

    void sadd(const std::string& key, const std::string& value);

    void srem(const std::string& key, const std::string& value);

    long scard(const std::string& key);

    bool sismember(const std::string& key, const std::string& value);

    std::string srandmember(const std::string& key);

    std::set<std::string> smembers(const std::string& key);

    std::set<std::string> sdiff(const std::string& keyA, const std::string& keyB);

    std::set<std::string> sinter(const std::string& keyA, const std::string& keyB);

    int lpushLong(const std::string& key, long value);

    int rpushLong(const std::string& key, long value);

    void flushAll();

    long listLength(const std::string& queueName);

    long getTtl(const std::string& key);

    void expire(const std::string& key, long expiry);

    void refresh();

    // ---- Locks ----
    uint32_t acquireLock(const std::string& key, int expirySeconds);

    void releaseLock(const std::string& key, uint32_t lockId);

    void delIfEq(const std::string& key, uint32_t value);

    bool setnxex(const std::string& key, long value, int expirySeconds);

    long getLong(const std::string& key);

    void setLong(const std::string& key, long value);

    // ---- Queueing ----
    void enqueue(const std::string& queueName, const std::string& value);

    void enqueueBytes(const std::string& queueName, const std::vector<uint8_t>& value);

    void enqueueBytes(const std::string& queueName, const uint8_t* buffer, size_t bufferLen);

    std::string dequeue(const std::string& queueName, int timeout = 60000);

    std::vector<uint8_t> dequeueBytes(const std::string& queueName, int timeout = 60000);

    void dequeueBytes(const std::string& queueName, uint8_t* buffer, size_t bufferLen, int timeout = 60000);

    void dequeueMultiple(const std::string& queueName, uint8_t* buff, long buffLen, long nElems);

    // ---- Scheduler notification ----
    void publishSchedulerResult(const std::string& key, const std::string& statusKey, const std::vector<uint8_t>& result);

  private:
    explicit Redis(RedisRole roleIn);

    RedisRole role;
    std::mutex mx;
    std::condition_variable listCv;
    std::unordered_map<std::string, std::vector<uint8_t>> strings;
    std::unordered_map<std::string, std::set<std::string>> sets;
    std::unordered_map<std::string, std::deque<std::vector<uint8_t>>> lists;
    std::unordered_map<std::string, long> expiries; // epoch ms
    uint32_t nextLockId = 1;

    bool isExpiredLocked(const std::string& key);

    std::vector<uint8_t> popFront(const std::string& queueName, int timeoutMs);
};

}



// ==========================================================================
// state/InMemoryStateRegistry.h
// ==========================================================================
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

// ==========================================================================
// state/StateKeyValue.h
// ==========================================================================
// One state value: host shared memory that is lazily claimed, pulled from /
// pushed to the key's main host in chunks, with per-byte dirty and pulled
// masks (reference: include/faabric/state/StateKeyValue.h:16-175,
// src/state/StateKeyValue.cpp:17-629).
//
// GPU extension: a value can have a *device-resident* copy (HBM) next to the
// host copy.  getDevicePtr() lazily uploads it; host writes invalidate chunks
// of it and device writes are brought back with syncFromDevice() (host-pinned
// staging, chunked copies on a dedicated stream).



#define STATE_STREAMING_CHUNK_SIZE (64 * 1024)

namespace faabric::state {

// A chunk of a state value (owned copy of the bytes)
class StateChunk
{
  public:
    StateChunk(long offsetIn, size_t lengthIn, uint8_t* dataIn)
      : offset(offsetIn)
      , length(lengthIn)
      , data(dataIn)
    {}

    StateChunk(long offsetIn, std::vector<uint8_t>& dataIn)
      : offset(offsetIn)
      , length(dataIn.size())
      , data(dataIn.data())
    {}

    long offset;
    size_t length;
    // Note - this pointer will always refer to chunks of the underlying
    // state, so does not need to be deleted
    uint8_t* data;
};

class StateKeyValueException : public faabric::util::FaabricException
{
  public:
    explicit StateKeyValueException(const std::string& message)
      : FaabricException(message)
    {}
};

// Remote lock acquisition: time per attempt and attempts before giving up
constexpr int REMOTE_LOCK_TIMEOUT_SECS(1);
constexpr int REMOTE_LOCK_MAX_RETRIES(100);

class StateKeyValue
{
  public:
    StateKeyValue(const std::string& userIn, const std::string& keyIn, size_t sizeIn);

    StateKeyValue(const std::string& userIn, const std::string& keyIn);

    virtual ~StateKeyValue();

    const std::string user;

    const std::string key;

    static uint32_t waitOnRedisRemoteLock(const std::string& redisKey);

    void get(uint8_t* buffer);

    uint8_t* get();

    void getChunk(long offset, uint8_t* buffer, size_t length);

    uint8_t* getChunk(long offset, long len);

    std::vector<StateChunk> getAllChunks();

    void set(const uint8_t* buffer);

    void setChunk(long offset, const uint8_t* buffer, size_t length);

    void append(const uint8_t* buffer, size_t length);

    void getAppended(uint8_t* buffer, size_t length, long nValues);

    void clearAppended();

    void mapSharedMemory(void* destination, long pagesOffset, long nPages);

    void unmapSharedMemory(void* mappedAddr);

    // Unmaps exactly the `nPages` that were mapped (the one-argument form
    // keeps the reference's behaviour: it unmaps the size of the whole value)
    void unmapSharedMemory(void* mappedAddr, long nPages);

    void flagDirty();

    void flagChunkDirty(long offset, long len);

    size_t size() const;

    size_t getSharedMemorySize() const;

    void pull();

    void pushFull();

    void pushPartial();

    void pushPartialMask(const std::shared_ptr<StateKeyValue>& maskKv);

    void lockRead();

    void unlockRead();

    void lockWrite();

    void unlockWrite();

    // ---- device-resident copy ----
    // Device pointer of the value on `device` (uploaded on first use; chunks
    // the host has written since are refreshed).  nullptr without a GPU.
    uint8_t* getDevicePtr(int device, void* stream = nullptr);

    // Mark a range as modified on the device
    void flagDeviceChunkDirty(long offset, long len);

    // Copy device-dirty chunks back to the host copy and flag them dirty so a
    // later pushPartial ships them
    void syncFromDevice(void* stream = nullptr);

    bool hasDeviceCopy() const { return deviceCopy.valid(); }

  protected:
    bool fullyAllocated = false;
    std::shared_mutex valueMutex;

    size_t valueSize = 0;
    size_t sharedMemSize = 0;
    void* sharedMemory = nullptr;

    void doSet(const uint8_t* data);

    void doSetChunk(long offset, const uint8_t* buffer, size_t length);

    void doPullChunk(bool lazy, long offset, size_t length);

    void doPushPartial(const uint8_t* dirtyMaskBytes);

    void configureSize();

    void checkSizeConfigured();

    void markDirtyChunk(long offset, long len);

    bool isChunkPulled(long offset, size_t length);

    void allocateChunk(long offset, size_t length);

    void reserveStorage();

    std::vector<StateChunk> getDirtyChunks(const uint8_t* dirtyMaskBytes);

    void zeroDirtyMask();

    // ---- backend hooks ----
    // Size of the authoritative copy (0 if unknown); lets a size-less replica
    // configure itself on first use
    virtual size_t sizeFromRemote() { return 0; }

    virtual void pullFromRemote() = 0;

    virtual void pullChunkFromRemote(long offset, size_t length) = 0;

    virtual void pushToRemote() = 0;

    virtual void pushPartialToRemote(const std::vector<StateChunk>& dirtyChunks) = 0;

    virtual void appendToRemote(const uint8_t* data, size_t length) = 0;

    virtual void pullAppendedFromRemote(uint8_t* data, size_t length, long nValues) = 0;

    virtual void clearAppendedFromRemote() = 0;

    void doPull(bool lazy);

  private:
    std::atomic<bool> isDirty = false;
    std::vector<uint8_t> dirtyMask;
    std::vector<uint8_t> pulledMask;

    // Device copy + which host chunks are newer than it / device-dirty chunks
    faabric::util::DeviceRegion deviceCopy;
    int deviceId = -1;
    bool hostRegistered = false;
    std::vector<uint8_t> hostNewerChunks;   // per STATE_STREAMING_CHUNK
    std::vector<uint8_t> deviceDirtyChunks; // per STATE_STREAMING_CHUNK
    void invalidateDeviceRange(long offset, long len);
};

}

// ==========================================================================
// state/InMemoryStateKeyValue.h
// ==========================================================================
namespace faabric::state {

enum InMemoryStateKeyStatus
{
    NOT_MASTER,
    MASTER,
};

class AppendedInMemoryState
{
  public:
    AppendedInMemoryState(size_t lengthIn, std::unique_ptr<uint8_t[]>&& dataIn)
      : length(lengthIn)
      , data(std::move(dataIn))
    {}

    size_t length;
    std::unique_ptr<uint8_t[]> data;
};

// Main host holds the authoritative bytes; other hosts pull / push chunks
// through StateClient (reference: src/state/InMemoryStateKeyValue.cpp:15-186)
class InMemoryStateKeyValue final : public StateKeyValue
{
  public:
    InMemoryStateKeyValue(const std::string& userIn,
                          const std::string& keyIn,
                          size_t sizeIn,
                          const std::string& thisIPIn);

    InMemoryStateKeyValue(const std::string& userIn,
                          const std::string& keyIn,
                          const std::string& thisIPIn);

    static size_t getStateSizeFromRemote(const std::string& userIn,
                                         const std::string& keyIn,
                                         const std::string& thisIPIn);

    static void deleteFromRemote(const std::string& userIn,
                                 const std::string& keyIn,
                                 const std::string& thisIPIn);

    static void clearAll(bool global);

    bool isMaster();

    AppendedInMemoryState& getAppendedValue(uint idx);

    // Exposed for the StateServer (it operates on the main copy)
    std::vector<AppendedInMemoryState>& getAppendedValues() { return appendedData; }

    std::mutex& getAppendedMutex() { return appendedMx; }

  private:
    const std::string thisIP;
    const std::string mainIP;
    InMemoryStateKeyStatus status;

    InMemoryStateRegistry& stateRegistry;

    std::mutex appendedMx;
    std::vector<AppendedInMemoryState> appendedData;

    size_t sizeFromRemote() override;

    void pullFromRemote() override;

    void pullChunkFromRemote(long offset, size_t length) override;

    void pushToRemote() override;

    void pushPartialToRemote(const std::vector<StateChunk>& dirtyChunks) override;

    void appendToRemote(const uint8_t* data, size_t length) override;

    void pullAppendedFromRemote(uint8_t* data, size_t length, long nValues) override;

    void clearAppendedFromRemote() override;
};

}

// ==========================================================================
// state/RedisStateKeyValue.h
// ==========================================================================
namespace faabric::state {

// STATE_MODE=redis: the value lives in the Redis-compatible store
// (reference: src/state/RedisStateKeyValue.cpp:15-129)
class RedisStateKeyValue final : public StateKeyValue
{
  public:
    RedisStateKeyValue(const std::string& userIn, const std::string& keyIn, size_t sizeIn);

    RedisStateKeyValue(const std::string& userIn, const std::string& keyIn);

    static size_t getStateSizeFromRemote(const std::string& userIn, const std::string& keyIn);

    static void deleteFromRemote(const std::string& userIn, const std::string& keyIn);

    static void clearAll(bool global);

  private:
    const std::string joinedKey;

    size_t sizeFromRemote() override;

    void pullFromRemote() override;

    void pullChunkFromRemote(long offset, size_t length) override;

    void pushToRemote() override;

    void pushPartialToRemote(const std::vector<StateChunk>& dirtyChunks) override;

    void appendToRemote(const uint8_t* data, size_t length) override;

    void pullAppendedFromRemote(uint8_t* data, size_t length, long nValues) override;

    void clearAppendedFromRemote() override;
};

}

// ==========================================================================
// state/DeviceStateKeyValue.h
// ==========================================================================
namespace faabric::state {

struct DeviceStateRun
{
    uint64_t offset;
    uint64_t length;
};

// A state value whose authoritative bytes live in HBM of the GPU that first
// claimed the key (the "main" GPU, the role of the main host in the
// reference).  Other GPUs hold replicas and move data with device copies over
// NVLink: lazy chunk pulls are peer copies, a partial push is ONE kernel that
// scans a device-resident dirty mask (one byte per 128-byte block), stores the
// dirty blocks straight into the main copy and clears the mask.  The host sees
// the value through a pinned mirror that is filled on demand.
// (Reference counterpart: StateKeyValue pull / push / dirty-chunk scan,
// src/state/StateKeyValue.cpp:61,394-543,592-629.)
class DeviceStateKeyValue
{
  public:
    DeviceStateKeyValue(std::string userIn,
                        std::string keyIn,
                        size_t sizeIn,
                        int deviceIn,
                        std::shared_ptr<DeviceStateKeyValue> mainIn);

    ~DeviceStateKeyValue();

    const std::string user;
    const std::string key;

    size_t size() const { return valueSize; }

    int getDevice() const { return device; }

    bool isMain() const { return main == nullptr; }

    // Device pointer of this GPU's copy (the value itself on the main GPU)
    uint8_t* getDevicePtr() { return data.ptr; }

    // ---- replica <- main ----
    void pull(void* stream = nullptr);

    // Lazy: copies only the 64 KiB chunks of [offset, offset+len) this replica
    // has not pulled yet
    void pullChunk(long offset, size_t length, void* stream = nullptr);

    bool isChunkPulled(long offset, size_t length);

    // ---- replica -> main ----
    // Mark [offset, offset+len) as written on this GPU (device-resident mask;
    // user kernels may also set mask bytes themselves: getDirtyMaskPtr())
    void flagChunkDirty(long offset, long len, void* stream = nullptr);

    void flagDirty(void* stream = nullptr);

    uint8_t* getDirtyMaskPtr() { return mask.ptr; }

    // Fused dirty scan + push + clear.  Returns the number of bytes pushed
    // (block granularity); synchronises `stream`.
    uint64_t pushPartial(void* stream = nullptr);

    void pushFull(void* stream = nullptr);

    // Dirty runs (block granularity) without pushing: device scan kernel
    std::vector<DeviceStateRun> getDirtyChunks(void* stream = nullptr);

    // ---- host access (pinned mirror, filled / flushed on demand) ----
    void get(uint8_t* buffer);

    void getChunk(long offset, uint8_t* buffer, size_t length);

    void set(const uint8_t* buffer);

    void setChunk(long offset, const uint8_t* buffer, size_t length);

    // Pinned host view of the whole value as of now (D2H copy of what changed
    // since the last call is not tracked: the whole value is refreshed)
    uint8_t* syncHostMirror(void* stream = nullptr);

    uint64_t getPushKernelLaunches() const { return pushLaunches; }

    uint64_t getBytesPulled() const { return bytesPulled; }

  private:
    size_t valueSize;
    int device;
    std::shared_ptr<DeviceStateKeyValue> main; // null on the main copy
    faabric::util::DeviceRegion data;
    faabric::util::DeviceRegion mask;  // one byte per FB_STATE_BLOCK_BYTES
    faabric::util::DeviceRegion stats; // 16 bytes
    faabric::util::DeviceRegion hostMirror; // pinned, lazy
    std::vector<uint8_t> pulledChunks; // per STATE_STREAMING_CHUNK_SIZE
    std::mutex mx;
    uint64_t pushLaunches = 0;
    uint64_t bytesPulled = 0;

    void checkRange(long offset, size_t length) const;
};

}

// ==========================================================================
// state/State.h
// ==========================================================================
#define STATE_INPROC_LABEL_KV "state-kv"

namespace faabric::state {

enum StateCalls
{
    NoStateCall = 0,
    Pull = 1,
    Push = 2,
    Size = 3,
    Append = 4,
    ClearAppended = 5,
    PullAppended = 6,
    Delete = 7,
};

// Process-wide registry of key-values (reference: src/state/State.cpp:14-183)
class State
{
  public:
    explicit State(std::string thisIPIn);

    size_t getStateSize(const std::string& user, const std::string& keyIn);

    std::shared_ptr<StateKeyValue> getKV(const std::string& user,
                                         const std::string& key,
                                         size_t size);

    std::shared_ptr<StateKeyValue> getKV(const std::string& user,
                                         const std::string& key);

    void forceClearAll(bool global);

    void deleteKV(const std::string& userIn, const std::string& keyIn);

    void deleteKVLocally(const std::string& userIn, const std::string& keyIn);

    size_t getKVCount();

    // ---- device-resident values ----
    // The first GPU to ask for user/key becomes its main GPU; later callers on
    // other GPUs get replicas that talk to it over NVLink.  One object per
    // (key, device).
    std::shared_ptr<DeviceStateKeyValue> getDeviceKV(const std::string& user,
                                                    const std::string& key,
                                                    size_t size,
                                                    int device);

    void deleteDeviceKV(const std::string& user, const std::string& key);

    size_t getDeviceKVCount();

    std::string getThisIP();

  private:
    const std::string thisIP;

    std::unordered_map<std::string, std::shared_ptr<StateKeyValue>> kvMap;
    std::shared_mutex mapMutex;
    // "user_key" -> per-device copies; entry -1 names the main device
    std::unordered_map<std::string, std::map<int, std::shared_ptr<DeviceStateKeyValue>>> deviceKvMap;
    std::unordered_map<std::string, int> deviceKvMain;

    std::shared_ptr<StateKeyValue> doGetKV(const std::string& user,
                                           const std::string& key,
                                           bool sizeless,
                                           size_t size);
};

State& getGlobalState();

}

// ==========================================================================
// state/StateClient.h
// ==========================================================================
namespace faabric::state {

// One synchronous RPC per 64 KiB chunk (reference: src/state/StateClient.cpp)
class StateClient : public faabric::transport::MessageEndpointClient
{
  public:
    explicit StateClient(const std::string& userIn,
                         const std::string& keyIn,
                         const std::string& hostIn);

    const std::string user;
    const std::string key;

    void pushChunks(const std::vector<StateChunk>& chunks);

    void pullChunks(const std::vector<StateChunk>& chunks, uint8_t* bufferStart);

    void append(const uint8_t* data, size_t length);

    void pullAppended(uint8_t* buffer, size_t length, long nValues);

    void clearAppended();

    size_t stateSize();

    void deleteState();

    void lock();

    void unlock();

  private:
    void sendStateRequest(faabric::state::StateCalls header, const uint8_t* data, int length);

    void logRequest(const std::string& op);
};

}

// ==========================================================================
// state/StateServer.h
// ==========================================================================
namespace faabric::state {

class StateServer final : public faabric::transport::MessageEndpointServer
{
  public:
    explicit StateServer(State& stateIn);

  private:
    State& state;

    void logOperation(const std::string& op);

    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    // Sync methods
    std::string recvSize(std::span<const uint8_t> buffer);

    std::string recvPull(std::span<const uint8_t> buffer);

    std::string recvPush(std::span<const uint8_t> buffer);

    std::string recvAppend(std::span<const uint8_t> buffer);

    std::string recvPullAppended(std::span<const uint8_t> buffer);

    std::string recvClearAppended(std::span<const uint8_t> buffer);

    std::string recvDelete(std::span<const uint8_t> buffer);
};

}

