// One state value: host shared memory that is lazily claimed, pulled from /
// pushed to the key's main host in chunks, with per-byte dirty and pulled
// masks (reference: include/faabric/state/StateKeyValue.h:16-175,
// src/state/StateKeyValue.cpp:17-629).
//
// GPU extension: a value can have a *device-resident* copy (HBM) next to the
// host copy.  getDevicePtr() lazily uploads it; host writes invalidate chunks
// of it and device writes are brought back with syncFromDevice() (host-pinned
// staging, chunked copies on a dedicated stream).
#pragma once

#include <faabric/util/exception.h>
#include <faabric/util/memory.h>

#include <atomic>
#include <cstdint>
#include <memory>
#include <set>
#include <shared_mutex>
#include <string>
#include <vector>

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
