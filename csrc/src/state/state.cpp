// State layer: registry, key-values (host shared memory + optional device copy),
// in-memory backend with main-host election, store-backed backend, RPC.
#include <faabric/planner/PlannerClient.h>
#include <faabric/redis/Redis.h>
#include <faabric/state/InMemoryStateKeyValue.h>
#include <faabric/state/InMemoryStateRegistry.h>
#include <faabric/state/RedisStateKeyValue.h>
#include <faabric/state/State.h>
#include <faabric/state/StateClient.h>
#include <faabric/state/StateKeyValue.h>
#include <faabric/state/StateServer.h>
#include <faabric/transport/common.h>
#include <faabric/util/bytes.h>
#include <faabric/util/macros.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/memory.h>
#include <faabric/util/timing.h>

#include <cuda_runtime.h>

#include <cstring>
#include <sys/mman.h>
#include <thread>

#define ONES_BITMASK 0b11111111
#define ZERO_BITMASK 0b00000000
#define MAIN_KEY_PREFIX "main_"

namespace faabric::state {

static std::string keyFor(const std::string& user, const std::string& key)
{
    return faabric::util::keyForUser(user, key);
}

// ---------------------------------------------------------------------------
// State registry
// ---------------------------------------------------------------------------
State& getGlobalState()
{
    static State s(faabric::transport::getThisHostAddress());
    return s;
}

State::State(std::string thisIPIn)
  : thisIP(std::move(thisIPIn))
{}

std::string State::getThisIP()
{
    return thisIP;
}

void State::forceClearAll(bool global)
{
    auto& conf = faabric::util::getSystemConfig();
    if (conf.stateMode == "redis") {
        RedisStateKeyValue::clearAll(global);
    } else {
        InMemoryStateKeyValue::clearAll(global);
    }
    std::unique_lock<std::shared_mutex> lock(mapMutex);
    kvMap.clear();
    deviceKvMap.clear();
    deviceKvMain.clear();
}

size_t State::getStateSize(const std::string& user, const std::string& keyIn)
{
    if (user.empty() || keyIn.empty()) {
        throw std::runtime_error("Attempting to get state size with empty user or key");
    }
    std::string lookup = keyFor(user, keyIn);
    {
        std::shared_lock<std::shared_mutex> lock(mapMutex);
        auto it = kvMap.find(lookup);
        if (it != kvMap.end()) {
            return it->second->size();
        }
    }
    auto& conf = faabric::util::getSystemConfig();
    if (conf.stateMode == "redis") {
        return RedisStateKeyValue::getStateSizeFromRemote(user, keyIn);
    }
    if (conf.stateMode == "inmemory") {
        return InMemoryStateKeyValue::getStateSizeFromRemote(user, keyIn, thisIP);
    }
    throw std::runtime_error("Unrecognised state mode: " + conf.stateMode);
}

void State::deleteKV(const std::string& userIn, const std::string& keyIn)
{
    auto& conf = faabric::util::getSystemConfig();
    if (conf.stateMode == "redis") {
        RedisStateKeyValue::deleteFromRemote(userIn, keyIn);
    } else if (conf.stateMode == "inmemory") {
        InMemoryStateKeyValue::deleteFromRemote(userIn, keyIn, thisIP);
    } else {
        throw std::runtime_error("Unrecognised state mode: " + conf.stateMode);
    }
    deleteKVLocally(userIn, keyIn);
}

void State::deleteKVLocally(const std::string& userIn, const std::string& keyIn)
{
    std::unique_lock<std::shared_mutex> lock(mapMutex);
    kvMap.erase(keyFor(userIn, keyIn));
}

std::shared_ptr<StateKeyValue> State::getKV(const std::string& user, const std::string& key)
{
    return doGetKV(user, key, true, 0);
}

std::shared_ptr<StateKeyValue> State::getKV(const std::string& user, const std::string& key, size_t size)
{
    return doGetKV(user, key, false, size);
}

std::shared_ptr<StateKeyValue> State::doGetKV(const std::string& user,
                                              const std::string& key,
                                              bool sizeless,
                                              size_t size)
{
    if (user.empty() || key.empty()) {
        throw std::runtime_error("Attempting to access state with empty user or key (" + user + "/" + key + ")");
    }
    std::string lookup = keyFor(user, key);
    {
        std::shared_lock<std::shared_mutex> lock(mapMutex);
        auto it = kvMap.find(lookup);
        if (it != kvMap.end()) {
            return it->second;
        }
    }
    std::unique_lock<std::shared_mutex> lock(mapMutex);
    auto it = kvMap.find(lookup);
    if (it != kvMap.end()) {
        return it->second;
    }
    auto& conf = faabric::util::getSystemConfig();
    std::shared_ptr<StateKeyValue> kv;
    if (conf.stateMode == "redis") {
        kv = sizeless ? std::make_shared<RedisStateKeyValue>(user, key)
                      : std::make_shared<RedisStateKeyValue>(user, key, size);
    } else if (conf.stateMode == "inmemory") {
        kv = sizeless ? std::make_shared<InMemoryStateKeyValue>(user, key, thisIP)
                      : std::make_shared<InMemoryStateKeyValue>(user, key, size, thisIP);
    } else {
        throw std::runtime_error("Unrecognised state mode: " + conf.stateMode);
    }
    kvMap.emplace(lookup, kv);
    return kv;
}

size_t State::getKVCount()
{
    std::shared_lock<std::shared_mutex> lock(mapMutex);
    return kvMap.size();
}

// ---------------------------------------------------------------------------
// StateKeyValue
// ---------------------------------------------------------------------------
StateKeyValue::StateKeyValue(const std::string& userIn, const std::string& keyIn)
  : StateKeyValue(userIn, keyIn, 0)
{}

StateKeyValue::StateKeyValue(const std::string& userIn, const std::string& keyIn, size_t sizeIn)
  : user(userIn)
  , key(keyIn)
  , valueSize(sizeIn)
{
    if (sizeIn > 0) {
        configureSize();
    }
}

StateKeyValue::~StateKeyValue()
{
    if (hostRegistered && sharedMemory != nullptr) {
        cudaHostUnregister(sharedMemory);
        cudaGetLastError();
    }
    if (sharedMemory != nullptr) {
        ::munmap(sharedMemory, sharedMemSize);
        sharedMemory = nullptr;
    }
}

void StateKeyValue::configureSize()
{
    // Storage is page-rounded so it can be mapped into other address spaces
    size_t nPages = faabric::util::getRequiredHostPages(valueSize);
    sharedMemSize = nPages * faabric::util::HOST_PAGE_SIZE;
    sharedMemory = nullptr;
    dirtyMask.assign(valueSize, ZERO_BITMASK);
    pulledMask.assign(valueSize, ZERO_BITMASK);
    size_t nChunks = (valueSize + STATE_STREAMING_CHUNK_SIZE - 1) / STATE_STREAMING_CHUNK_SIZE;
    hostNewerChunks.assign(nChunks, 1);
    deviceDirtyChunks.assign(nChunks, 0);
}

void StateKeyValue::checkSizeConfigured()
{
    if (valueSize <= 0) {
        // Created without a size: ask whoever holds the value
        size_t remote = sizeFromRemote();
        if (remote > 0) {
            valueSize = remote;
            configureSize();
            return;
        }
        throw StateKeyValueException(std::string("State value size not set for ") + user + "/" + key);
    }
}

size_t StateKeyValue::size() const
{
    return valueSize;
}

size_t StateKeyValue::getSharedMemorySize() const
{
    return sharedMemSize;
}

void StateKeyValue::reserveStorage()
{
    checkSizeConfigured();
    if (sharedMemSize == 0) {
        throw StateKeyValueException("Reserving storage with no size for " + key);
    }
    // Reserve the whole range, claim pages on demand
    void* p = ::mmap(nullptr, sharedMemSize, PROT_NONE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        throw std::runtime_error("Failed mapping memory for KV");
    }
    sharedMemory = p;
}

void StateKeyValue::allocateChunk(long offset, size_t length)
{
    if (fullyAllocated) {
        return;
    }
    if (sharedMemory == nullptr) {
        reserveStorage();
    }
    auto chunk = faabric::util::getPageAlignedChunk(offset, (long)length);
    if ((size_t)(chunk.nBytesOffset + chunk.nBytesLength) > sharedMemSize) {
        throw StateKeyValueException("Allocating chunk beyond reserved storage for " + key);
    }
    uint8_t* base = BYTES(sharedMemory) + chunk.nBytesOffset;
    if (::mprotect(base, (size_t)chunk.nBytesLength, PROT_READ | PROT_WRITE) != 0) {
        throw std::runtime_error("Failed to claim pages for KV");
    }
    if (chunk.nBytesOffset == 0 && (size_t)chunk.nBytesLength >= sharedMemSize) {
        fullyAllocated = true;
    }
}

void StateKeyValue::zeroDirtyMask()
{
    std::fill(dirtyMask.begin(), dirtyMask.end(), ZERO_BITMASK);
}

void StateKeyValue::markDirtyChunk(long offset, long len)
{
    isDirty = true;
    std::fill(dirtyMask.begin() + offset, dirtyMask.begin() + offset + len, ONES_BITMASK);
}

bool StateKeyValue::isChunkPulled(long offset, size_t length)
{
    if (pulledMask.empty()) {
        return false;
    }
    for (size_t i = (size_t)offset; i < (size_t)offset + length; i++) {
        if (pulledMask[i] == 0) {
            return false;
        }
    }
    return true;
}

void StateKeyValue::invalidateDeviceRange(long offset, long len)
{
    if (hostNewerChunks.empty() || len <= 0) {
        return;
    }
    size_t first = (size_t)offset / STATE_STREAMING_CHUNK_SIZE;
    size_t last = (size_t)(offset + len - 1) / STATE_STREAMING_CHUNK_SIZE;
    for (size_t c = first; c <= last && c < hostNewerChunks.size(); c++) {
        hostNewerChunks[c] = 1;
    }
}

void StateKeyValue::doPull(bool lazy)
{
    // Resolve the size first: a size-less replica learns it here
    checkSizeConfigured();
    doPullChunk(lazy, 0, valueSize);
}

void StateKeyValue::pull()
{
    doPull(false);
}

void StateKeyValue::doPullChunk(bool lazy, long offset, size_t length)
{
    PROF_START(statePullChunk)
    checkSizeConfigured();
    if ((size_t)offset + length > valueSize) {
        throw StateKeyValueException("Pulling chunk out of bounds of " + key);
    }
    // Lazy pulls only go remote for bytes we have never seen
    if (lazy && isChunkPulled(offset, length)) {
        return;
    }
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (lazy && isChunkPulled(offset, length)) {
        return;
    }
    allocateChunk(offset, length);
    if (offset == 0 && length == valueSize) {
        pullFromRemote();
    } else {
        pullChunkFromRemote(offset, length);
    }
    std::fill(pulledMask.begin() + offset, pulledMask.begin() + offset + length, ONES_BITMASK);
    invalidateDeviceRange(offset, (long)length);
    PROF_END(statePullChunk)
}

void StateKeyValue::get(uint8_t* buffer)
{
    doPull(true);
    std::shared_lock<std::shared_mutex> lock(valueMutex);
    memcpy(buffer, sharedMemory, valueSize);
}

uint8_t* StateKeyValue::get()
{
    doPull(true);
    return BYTES(sharedMemory);
}

void StateKeyValue::getChunk(long offset, uint8_t* buffer, size_t length)
{
    doPullChunk(true, offset, length);
    std::shared_lock<std::shared_mutex> lock(valueMutex);
    memcpy(buffer, BYTES(sharedMemory) + offset, length);
}

uint8_t* StateKeyValue::getChunk(long offset, long len)
{
    doPullChunk(true, offset, (size_t)len);
    return BYTES(sharedMemory) + offset;
}

std::vector<StateChunk> StateKeyValue::getAllChunks()
{
    std::vector<StateChunk> chunks;
    uint8_t* base = BYTES(sharedMemory);
    for (size_t off = 0; off < valueSize; off += STATE_STREAMING_CHUNK_SIZE) {
        size_t len = std::min<size_t>(STATE_STREAMING_CHUNK_SIZE, valueSize - off);
        chunks.emplace_back((long)off, len, base + off);
    }
    return chunks;
}

void StateKeyValue::doSet(const uint8_t* buffer)
{
    checkSizeConfigured();
    if (sharedMemory == nullptr || !fullyAllocated) {
        allocateChunk(0, sharedMemSize);
    }
    memcpy(sharedMemory, buffer, valueSize);
}

void StateKeyValue::set(const uint8_t* buffer)
{
    checkSizeConfigured();
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    doSet(buffer);
    isDirty = true;
    markDirtyChunk(0, (long)valueSize);
    // A full write means nothing needs pulling any more
    std::fill(pulledMask.begin(), pulledMask.end(), ONES_BITMASK);
    invalidateDeviceRange(0, (long)valueSize);
}

void StateKeyValue::doSetChunk(long offset, const uint8_t* buffer, size_t length)
{
    checkSizeConfigured();
    // Writes may run into the page-rounded slack, not beyond it
    if ((size_t)offset + length > sharedMemSize) {
        throw StateKeyValueException("Setting state chunk too big for container. Key: " + key);
    }
    allocateChunk(offset, length);
    memcpy(BYTES(sharedMemory) + offset, buffer, length);
}

void StateKeyValue::setChunk(long offset, const uint8_t* buffer, size_t length)
{
    checkSizeConfigured();
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    doSetChunk(offset, buffer, length);
    size_t maskLen = std::min(length, valueSize > (size_t)offset ? valueSize - (size_t)offset : 0);
    if (maskLen > 0) {
        markDirtyChunk(offset, (long)maskLen);
        std::fill(pulledMask.begin() + offset, pulledMask.begin() + offset + maskLen, ONES_BITMASK);
    }
    invalidateDeviceRange(offset, (long)length);
}

void StateKeyValue::append(const uint8_t* buffer, size_t length)
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    appendToRemote(buffer, length);
}

void StateKeyValue::getAppended(uint8_t* buffer, size_t length, long nValues)
{
    std::shared_lock<std::shared_mutex> lock(valueMutex);
    pullAppendedFromRemote(buffer, length, nValues);
}

void StateKeyValue::clearAppended()
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    clearAppendedFromRemote();
}

void StateKeyValue::flagDirty()
{
    std::shared_lock<std::shared_mutex> lock(valueMutex);
    isDirty = true;
}

void StateKeyValue::flagChunkDirty(long offset, long len)
{
    checkSizeConfigured();
    std::shared_lock<std::shared_mutex> lock(valueMutex);
    markDirtyChunk(offset, len);
}

void StateKeyValue::mapSharedMemory(void* destination, long pagesOffset, long nPages)
{
    checkSizeConfigured();
    PROF_START(mapSharedMem)
    if (!faabric::util::isPageAligned(destination)) {
        SPDLOG_ERROR("Non-aligned destination for shared mapping of {}", key);
        throw std::runtime_error("Mapping misaligned shared memory");
    }
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    size_t offset = (size_t)pagesOffset * faabric::util::HOST_PAGE_SIZE;
    size_t length = (size_t)nPages * faabric::util::HOST_PAGE_SIZE;
    if (offset + length > sharedMemSize) {
        throw StateKeyValueException("Mapping shared memory beyond end of " + key);
    }
    allocateChunk((long)offset, length);
    // Alias the same physical pages at the caller's address (never pulls)
    void* res = ::mremap(BYTES(sharedMemory) + offset, 0, length, MREMAP_FIXED | MREMAP_MAYMOVE, destination);
    if (res == MAP_FAILED) {
        SPDLOG_ERROR("Failed mapping for {} at {} with size {}: {}", key, offset, length, strerror(errno));
        throw std::runtime_error("Failed mapping shared memory");
    }
    if (destination != res) {
        throw std::runtime_error("Misaligned shared memory mapping");
    }
    PROF_END(mapSharedMem)
}

void StateKeyValue::unmapSharedMemory(void* mappedAddr)
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!faabric::util::isPageAligned(mappedAddr)) {
        throw std::runtime_error("Attempting to unmap non-page-aligned memory");
    }
    if (::munmap(mappedAddr, sharedMemSize) != 0) {
        throw std::runtime_error("Failed unmapping shared memory");
    }
}

void StateKeyValue::unmapSharedMemory(void* mappedAddr, long nPages)
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!faabric::util::isPageAligned(mappedAddr)) {
        throw std::runtime_error("Attempting to unmap non-page-aligned memory");
    }
    if (nPages <= 0 || ::munmap(mappedAddr, (size_t)nPages * faabric::util::HOST_PAGE_SIZE) != 0) {
        throw std::runtime_error("Failed unmapping shared memory");
    }
}

void StateKeyValue::pushFull()
{
    if (!isDirty) {
        return;
    }
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!isDirty) {
        return;
    }
    pushToRemote();
    isDirty = false;
    zeroDirtyMask();
}

std::vector<StateChunk> StateKeyValue::getDirtyChunks(const uint8_t* dirtyMaskBytes)
{
    // Maximal runs of dirty bytes
    std::vector<StateChunk> chunks;
    uint8_t* base = BYTES(sharedMemory);
    bool inRun = false;
    long runStart = 0;
    for (size_t i = 0; i < valueSize; i++) {
        bool dirty = dirtyMaskBytes[i] != 0;
        if (dirty && !inRun) {
            inRun = true;
            runStart = (long)i;
        } else if (!dirty && inRun) {
            chunks.emplace_back(runStart, i - (size_t)runStart, base + runStart);
            inRun = false;
        }
    }
    if (inRun) {
        chunks.emplace_back(runStart, valueSize - (size_t)runStart, base + runStart);
    }
    return chunks;
}

void StateKeyValue::doPushPartial(const uint8_t* dirtyMaskBytes)
{
    PROF_START(pushPartial)
    if (!isDirty) {
        return;
    }
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!isDirty) {
        return;
    }
    std::vector<StateChunk> chunks = getDirtyChunks(dirtyMaskBytes);
    zeroDirtyMask();
    pushPartialToRemote(chunks);
    isDirty = false;
    PROF_END(pushPartial)
}

void StateKeyValue::pushPartial()
{
    doPushPartial(dirtyMask.data());
}

void StateKeyValue::pushPartialMask(const std::shared_ptr<StateKeyValue>& maskKv)
{
    if (maskKv->size() != valueSize) {
        throw StateKeyValueException("Different sizes: mask=" + std::to_string(maskKv->size()) + " and value=" + std::to_string(valueSize));
    }
    doPushPartial(maskKv->get());
}

void StateKeyValue::lockRead()
{
    valueMutex.lock_shared();
}

void StateKeyValue::unlockRead()
{
    valueMutex.unlock_shared();
}

void StateKeyValue::lockWrite()
{
    valueMutex.lock();
}

void StateKeyValue::unlockWrite()
{
    valueMutex.unlock();
}

uint32_t StateKeyValue::waitOnRedisRemoteLock(const std::string& redisKey)
{
    auto& redis = faabric::redis::Redis::getState();
    uint32_t lockId = redis.acquireLock(redisKey, REMOTE_LOCK_TIMEOUT_SECS);
    unsigned int retries = 0;
    while (lockId == 0) {
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        lockId = redis.acquireLock(redisKey, REMOTE_LOCK_TIMEOUT_SECS);
        if (++retries >= REMOTE_LOCK_MAX_RETRIES * 100) {
            SPDLOG_ERROR("Timed out waiting for lock on {}", redisKey);
            break;
        }
    }
    return lockId;
}

// ---- device copy ----
uint8_t* StateKeyValue::getDevicePtr(int device, void* stream)
{
    checkSizeConfigured();
    // Make sure the host copy is complete first
    doPull(true);
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!deviceCopy.valid()) {
        try {
            deviceCopy = faabric::util::allocateDeviceMemory(sharedMemSize, device);
        } catch (const std::exception& e) {
            SPDLOG_DEBUG("No device copy for {}: {}", key, e.what());
            return nullptr;
        }
        deviceId = device;
        std::fill(hostNewerChunks.begin(), hostNewerChunks.end(), 1);
    }
    if (sharedMemory == nullptr || !fullyAllocated) {
        allocateChunk(0, sharedMemSize);
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(deviceId);
    if (!hostRegistered) {
        // Pin the host value so chunk uploads / downloads are true async DMA
        if (cudaHostRegister(sharedMemory, sharedMemSize, cudaHostRegisterPortable) == cudaSuccess) {
            hostRegistered = true;
        } else {
            cudaGetLastError();
        }
    }
    // Refresh only what the host changed since the last upload
    for (size_t c = 0; c < hostNewerChunks.size(); c++) {
        if (!hostNewerChunks[c]) {
            continue;
        }
        size_t off = c * STATE_STREAMING_CHUNK_SIZE;
        size_t len = std::min<size_t>(STATE_STREAMING_CHUNK_SIZE, valueSize - off);
        cudaMemcpyAsync(deviceCopy.ptr + off, BYTES(sharedMemory) + off, len, cudaMemcpyHostToDevice, (cudaStream_t)stream);
        hostNewerChunks[c] = 0;
    }
    cudaStreamSynchronize((cudaStream_t)stream);
    if (prev >= 0) {
        cudaSetDevice(prev);
    }
    return deviceCopy.ptr;
}

void StateKeyValue::flagDeviceChunkDirty(long offset, long len)
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (deviceDirtyChunks.empty() || len <= 0) {
        return;
    }
    size_t first = (size_t)offset / STATE_STREAMING_CHUNK_SIZE;
    size_t last = (size_t)(offset + len - 1) / STATE_STREAMING_CHUNK_SIZE;
    for (size_t c = first; c <= last && c < deviceDirtyChunks.size(); c++) {
        deviceDirtyChunks[c] = 1;
    }
}

void StateKeyValue::syncFromDevice(void* stream)
{
    std::unique_lock<std::shared_mutex> lock(valueMutex);
    if (!deviceCopy.valid()) {
        return;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(deviceId);
    for (size_t c = 0; c < deviceDirtyChunks.size(); c++) {
        if (!deviceDirtyChunks[c]) {
            continue;
        }
        size_t off = c * STATE_STREAMING_CHUNK_SIZE;
        size_t len = std::min<size_t>(STATE_STREAMING_CHUNK_SIZE, valueSize - off);
        allocateChunk((long)off, len);
        cudaMemcpyAsync(BYTES(sharedMemory) + off, deviceCopy.ptr + off, len, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        markDirtyChunk((long)off, (long)len);
        std::fill(pulledMask.begin() + off, pulledMask.begin() + off + len, ONES_BITMASK);
        deviceDirtyChunks[c] = 0;
    }
    cudaStreamSynchronize((cudaStream_t)stream);
    if (prev >= 0) {
        cudaSetDevice(prev);
    }
}

// ---------------------------------------------------------------------------
// In-memory registry
// ---------------------------------------------------------------------------
InMemoryStateRegistry& getInMemoryStateRegistry()
{
    static InMemoryStateRegistry reg;
    return reg;
}

static std::string mainKeyFor(const std::string& user, const std::string& key)
{
    return MAIN_KEY_PREFIX + user + "_" + key;
}

std::string InMemoryStateRegistry::getMasterIP(const std::string& user,
                                               const std::string& key,
                                               const std::string& thisIP,
                                               bool claim)
{
    std::string lookup = keyFor(user, key);
    {
        std::shared_lock<std::shared_mutex> lock(mainMapMutex);
        auto it = mainMap.find(lookup);
        if (it != mainMap.end()) {
            return it->second;
        }
    }
    std::unique_lock<std::shared_mutex> lock(mainMapMutex);
    auto it = mainMap.find(lookup);
    if (it != mainMap.end()) {
        return it->second;
    }
    if (sharedViaPlanner) {
        std::string elected = faabric::planner::getPlannerClient().stateMain(user, key, thisIP, claim);
        if (elected.empty()) {
            throw StateKeyValueException("Found no main for state " + lookup);
        }
        mainMap[lookup] = elected;
        return elected;
    }
    // Consult (and possibly write) the shared store under its lock
    std::string storeKey = mainKeyFor(user, key);
    auto& redis = faabric::redis::Redis::getState();
    std::vector<uint8_t> mainBytes = redis.get(storeKey);
    if (mainBytes.empty() && !claim) {
        SPDLOG_TRACE("No main found for {}/{}", user, key);
        throw StateKeyValueException("Found no main for state " + storeKey);
    }
    if (mainBytes.empty()) {
        uint32_t lockId = StateKeyValue::waitOnRedisRemoteLock(storeKey);
        if (lockId == 0) {
            throw std::runtime_error("Unable to get remote lock for " + storeKey);
        }
        mainBytes = redis.get(storeKey);
        if (mainBytes.empty()) {
            mainBytes = faabric::util::stringToBytes(thisIP);
            redis.set(storeKey, mainBytes);
        }
        redis.releaseLock(storeKey, lockId);
    }
    std::string mainIP = faabric::util::bytesToString(mainBytes);
    mainMap[lookup] = mainIP;
    return mainIP;
}

std::string InMemoryStateRegistry::getMasterIPForOtherMaster(const std::string& userIn,
                                                             const std::string& keyIn,
                                                             const std::string& thisIP)
{
    std::string mainIP = getMasterIP(userIn, keyIn, thisIP, false);
    if (mainIP == thisIP) {
        throw std::runtime_error("Attempting to pull state size on main " + userIn + "/" + keyIn);
    }
    return mainIP;
}

void InMemoryStateRegistry::dropMain(const std::string& user, const std::string& key, bool everywhere)
{
    {
        std::unique_lock<std::shared_mutex> lock(mainMapMutex);
        mainMap.erase(keyFor(user, key));
    }
    if (!everywhere) {
        return;
    }
    if (sharedViaPlanner) {
        faabric::planner::getPlannerClient().stateMain(user, key, "", false, true);
    } else {
        faabric::redis::Redis::getState().del(mainKeyFor(user, key));
    }
}

void InMemoryStateRegistry::clear()
{
    std::unique_lock<std::shared_mutex> lock(mainMapMutex);
    mainMap.clear();
}

// ---------------------------------------------------------------------------
// In-memory KV
// ---------------------------------------------------------------------------
InMemoryStateKeyValue::InMemoryStateKeyValue(const std::string& userIn,
                                             const std::string& keyIn,
                                             size_t sizeIn,
                                             const std::string& thisIPIn)
  : StateKeyValue(userIn, keyIn, sizeIn)
  , thisIP(thisIPIn)
  , mainIP(getInMemoryStateRegistry().getMasterIP(user, key, thisIP, true))
  , status(thisIP == mainIP ? InMemoryStateKeyStatus::MASTER : InMemoryStateKeyStatus::NOT_MASTER)
  , stateRegistry(getInMemoryStateRegistry())
{
    SPDLOG_TRACE("Creating in-memory state key-value for {}/{} size {} (this host {}, main {})", user, key, sizeIn, thisIP, mainIP);
}

InMemoryStateKeyValue::InMemoryStateKeyValue(const std::string& userIn,
                                             const std::string& keyIn,
                                             const std::string& thisIPIn)
  : InMemoryStateKeyValue(userIn, keyIn, 0, thisIPIn)
{}

size_t InMemoryStateKeyValue::getStateSizeFromRemote(const std::string& userIn,
                                                     const std::string& keyIn,
                                                     const std::string& thisIPIn)
{
    std::string mainIP;
    try {
        mainIP = getInMemoryStateRegistry().getMasterIPForOtherMaster(userIn, keyIn, thisIPIn);
    } catch (const StateKeyValueException&) {
        return 0;
    }
    StateClient client(userIn, keyIn, mainIP);
    return client.stateSize();
}

void InMemoryStateKeyValue::deleteFromRemote(const std::string& userIn,
                                             const std::string& keyIn,
                                             const std::string& thisIPIn)
{
    InMemoryStateRegistry& reg = getInMemoryStateRegistry();
    std::string mainIP = reg.getMasterIP(userIn, keyIn, thisIPIn, false);
    if (mainIP == thisIPIn) {
        // Nothing remote to delete, but the key is up for election again
        reg.dropMain(userIn, keyIn, true);
        return;
    }
    StateClient client(userIn, keyIn, mainIP);
    client.deleteState();
    // The main withdrew the registration; forget our cached copy of it
    reg.dropMain(userIn, keyIn, false);
}

void InMemoryStateKeyValue::clearAll(bool global)
{
    getInMemoryStateRegistry().clear();
    if (global) {
        faabric::redis::Redis::getState().flushAll();
    }
}

bool InMemoryStateKeyValue::isMaster()
{
    return status == InMemoryStateKeyStatus::MASTER;
}

AppendedInMemoryState& InMemoryStateKeyValue::getAppendedValue(uint idx)
{
    return appendedData.at(idx);
}

size_t InMemoryStateKeyValue::sizeFromRemote()
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        return 0;
    }
    StateClient client(user, key, mainIP);
    return client.stateSize();
}

void InMemoryStateKeyValue::pullFromRemote()
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        return;
    }
    std::vector<StateChunk> chunks = getAllChunks();
    StateClient client(user, key, mainIP);
    client.pullChunks(chunks, BYTES(sharedMemory));
}

void InMemoryStateKeyValue::pullChunkFromRemote(long offset, size_t length)
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        return;
    }
    uint8_t* chunkStart = BYTES(sharedMemory) + offset;
    std::vector<StateChunk> chunks = { StateChunk(offset, length, chunkStart) };
    StateClient client(user, key, mainIP);
    client.pullChunks(chunks, BYTES(sharedMemory));
}

void InMemoryStateKeyValue::pushToRemote()
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        return;
    }
    std::vector<StateChunk> allChunks = getAllChunks();
    StateClient client(user, key, mainIP);
    client.pushChunks(allChunks);
}

void InMemoryStateKeyValue::pushPartialToRemote(const std::vector<StateChunk>& chunks)
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        // Nothing to be done
        return;
    }
    StateClient client(user, key, mainIP);
    client.pushChunks(chunks);
}

void InMemoryStateKeyValue::appendToRemote(const uint8_t* data, size_t length)
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        // Keep our own copy of the appended bytes
        auto copy = std::make_unique<uint8_t[]>(length);
        memcpy(copy.get(), data, length);
        std::lock_guard<std::mutex> lk(appendedMx);
        appendedData.emplace_back(length, std::move(copy));
    } else {
        StateClient client(user, key, mainIP);
        client.append(data, length);
    }
}

void InMemoryStateKeyValue::pullAppendedFromRemote(uint8_t* data, size_t length, long nValues)
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        std::lock_guard<std::mutex> lk(appendedMx);
        if ((size_t)nValues > appendedData.size()) {
            SPDLOG_ERROR("Trying to read {} appended values, but only {} set", nValues, appendedData.size());
            throw std::runtime_error("Reading more appended values than exist");
        }
        size_t off = 0;
        for (long i = 0; i < nValues; i++) {
            AppendedInMemoryState& v = appendedData.at(i);
            if (off + v.length > length) {
                throw std::runtime_error("Buffer too small for appended values");
            }
            memcpy(data + off, v.data.get(), v.length);
            off += v.length;
        }
    } else {
        StateClient client(user, key, mainIP);
        client.pullAppended(data, length, nValues);
    }
}

void InMemoryStateKeyValue::clearAppendedFromRemote()
{
    if (status == InMemoryStateKeyStatus::MASTER) {
        std::lock_guard<std::mutex> lk(appendedMx);
        appendedData.clear();
    } else {
        StateClient client(user, key, mainIP);
        client.clearAppended();
    }
}

// ---------------------------------------------------------------------------
// Store-backed KV
// ---------------------------------------------------------------------------
RedisStateKeyValue::RedisStateKeyValue(const std::string& userIn, const std::string& keyIn, size_t sizeIn)
  : StateKeyValue(userIn, keyIn, sizeIn)
  , joinedKey(keyFor(userIn, keyIn))
{}

RedisStateKeyValue::RedisStateKeyValue(const std::string& userIn, const std::string& keyIn)
  : RedisStateKeyValue(userIn, keyIn, 0)
{}

size_t RedisStateKeyValue::getStateSizeFromRemote(const std::string& userIn, const std::string& keyIn)
{
    return faabric::redis::Redis::getState().strlen(keyFor(userIn, keyIn));
}

void RedisStateKeyValue::deleteFromRemote(const std::string& userIn, const std::string& keyIn)
{
    faabric::redis::Redis::getState().del(keyFor(userIn, keyIn));
}

void RedisStateKeyValue::clearAll(bool global)
{
    if (global) {
        faabric::redis::Redis::getState().flushAll();
    }
}

size_t RedisStateKeyValue::sizeFromRemote()
{
    return faabric::redis::Redis::getState().strlen(joinedKey);
}

void RedisStateKeyValue::pullFromRemote()
{
    PROF_START(statePull)
    faabric::redis::Redis::getState().get(joinedKey, BYTES(sharedMemory), valueSize);
    PROF_END(statePull)
}

void RedisStateKeyValue::pullChunkFromRemote(long offset, size_t length)
{
    PROF_START(stateChunkPull)
    faabric::redis::Redis::getState().getRange(joinedKey, BYTES(sharedMemory) + offset, length, offset, offset + (long)length - 1);
    PROF_END(stateChunkPull)
}

void RedisStateKeyValue::pushToRemote()
{
    PROF_START(pushFull)
    faabric::redis::Redis::getState().set(joinedKey, BYTES(sharedMemory), valueSize);
    PROF_END(pushFull)
}

void RedisStateKeyValue::pushPartialToRemote(const std::vector<StateChunk>& dirtyChunks)
{
    PROF_START(updatePipeline)
    auto& redis = faabric::redis::Redis::getState();
    for (const auto& c : dirtyChunks) {
        redis.setRangePipeline(joinedKey, c.offset, c.data, c.length);
    }
    redis.flushPipeline((long)dirtyChunks.size());
    PROF_END(updatePipeline)
}

void RedisStateKeyValue::appendToRemote(const uint8_t* data, size_t length)
{
    faabric::redis::Redis::getState().enqueueBytes(joinedKey + "_appended", data, length);
}

void RedisStateKeyValue::pullAppendedFromRemote(uint8_t* data, size_t length, long nValues)
{
    faabric::redis::Redis::getState().dequeueMultiple(joinedKey + "_appended", data, (long)length, nValues);
}

void RedisStateKeyValue::clearAppendedFromRemote()
{
    faabric::redis::Redis::getState().del(joinedKey + "_appended");
}

// ---------------------------------------------------------------------------
// Client
// ---------------------------------------------------------------------------
StateClient::StateClient(const std::string& userIn, const std::string& keyIn, const std::string& hostIn)
  : faabric::transport::MessageEndpointClient(hostIn, STATE_ASYNC_PORT, STATE_SYNC_PORT)
  , user(userIn)
  , key(keyIn)
{}

void StateClient::logRequest(const std::string& op)
{
    SPDLOG_TRACE("Requesting {} on {}/{} at {}", op, user, key, host);
}

void StateClient::sendStateRequest(faabric::state::StateCalls header, const uint8_t* data, int length)
{
    faabric::StateRequest request;
    request.set_user(user);
    request.set_key(key);
    if (length > 0) {
        request.set_data(data, (size_t)length);
    }
    faabric::EmptyResponse resp;
    syncSend(header, &request, &resp);
}

void StateClient::pushChunks(const std::vector<StateChunk>& chunks)
{
    logRequest("push-chunks");
    for (const auto& chunk : chunks) {
        // Stream big chunks in 64 KiB parts
        for (size_t done = 0; done < chunk.length; done += STATE_STREAMING_CHUNK_SIZE) {
            size_t len = std::min<size_t>(STATE_STREAMING_CHUNK_SIZE, chunk.length - done);
            faabric::StatePart part;
            part.set_user(user);
            part.set_key(key);
            part.set_offset((uint64_t)chunk.offset + done);
            part.set_data(chunk.data + done, len);
            faabric::EmptyResponse resp;
            syncSend(faabric::state::StateCalls::Push, &part, &resp);
        }
    }
}

void StateClient::pullChunks(const std::vector<StateChunk>& chunks, uint8_t* bufferStart)
{
    logRequest("pull-chunks");
    for (const auto& chunk : chunks) {
        for (size_t done = 0; done < chunk.length; done += STATE_STREAMING_CHUNK_SIZE) {
            size_t len = std::min<size_t>(STATE_STREAMING_CHUNK_SIZE, chunk.length - done);
            faabric::StateChunkRequest request;
            request.set_user(user);
            request.set_key(key);
            request.set_offset((uint64_t)chunk.offset + done);
            request.set_chunksize(len);
            faabric::StatePart response;
            syncSend(faabric::state::StateCalls::Pull, &request, &response);
            if (response.data().size() != len) {
                throw std::runtime_error("Pulled state chunk has the wrong size");
            }
            memcpy(bufferStart + response.offset(), response.data().data(), response.data().size());
        }
    }
}

void StateClient::append(const uint8_t* data, size_t length)
{
    logRequest("append");
    sendStateRequest(faabric::state::StateCalls::Append, data, (int)length);
}

void StateClient::pullAppended(uint8_t* buffer, size_t length, long nValues)
{
    logRequest("pull-appended");
    faabric::StateAppendedRequest request;
    request.set_user(user);
    request.set_key(key);
    request.set_nvalues((uint32_t)nValues);
    faabric::StateAppendedResponse response;
    syncSend(faabric::state::StateCalls::PullAppended, &request, &response);
    size_t off = 0;
    for (const auto& v : response.values()) {
        if (off + v.data().size() > length) {
            throw std::runtime_error("Buffer not large enough for appended data (offset=" + std::to_string(off) + ", length=" + std::to_string(length) + ")");
        }
        memcpy(buffer + off, v.data().data(), v.data().size());
        off += v.data().size();
    }
}

void StateClient::clearAppended()
{
    logRequest("clear-appended");
    sendStateRequest(faabric::state::StateCalls::ClearAppended, nullptr, 0);
}

size_t StateClient::stateSize()
{
    logRequest("state-size");
    faabric::StateRequest request;
    request.set_user(user);
    request.set_key(key);
    faabric::StateSizeResponse response;
    syncSend(faabric::state::StateCalls::Size, &request, &response);
    return response.statesize();
}

void StateClient::deleteState()
{
    logRequest("delete");
    sendStateRequest(faabric::state::StateCalls::Delete, nullptr, 0);
}

void StateClient::lock() {}

void StateClient::unlock() {}

// ---------------------------------------------------------------------------
// Server
// ---------------------------------------------------------------------------
#define KV_FROM_REQUEST(request)                                               \
    auto kv = std::static_pointer_cast<InMemoryStateKeyValue>(                 \
      state.getKV((request).user(), (request).key()));

StateServer::StateServer(State& stateIn)
  : faabric::transport::MessageEndpointServer(STATE_ASYNC_PORT,
                                              STATE_SYNC_PORT,
                                              STATE_INPROC_LABEL,
                                              faabric::util::getSystemConfig().stateServerThreads)
  , state(stateIn)
{}

void StateServer::logOperation(const std::string& op)
{
    SPDLOG_TRACE("Received {}", op);
}

void StateServer::doAsyncRecv(transport::Message& message)
{
    throw std::runtime_error("State server does not support async recv");
}

std::string StateServer::doSyncRecv(transport::Message& message)
{
    uint8_t header = message.getMessageCode();
    switch (header) {
        case faabric::state::StateCalls::Pull:
            return recvPull(message.udata());
        case faabric::state::StateCalls::Push:
            return recvPush(message.udata());
        case faabric::state::StateCalls::Size:
            return recvSize(message.udata());
        case faabric::state::StateCalls::Append:
            return recvAppend(message.udata());
        case faabric::state::StateCalls::ClearAppended:
            return recvClearAppended(message.udata());
        case faabric::state::StateCalls::PullAppended:
            return recvPullAppended(message.udata());
        case faabric::state::StateCalls::Delete:
            return recvDelete(message.udata());
        default:
            throw std::runtime_error("Unrecognized state call header: " + std::to_string(header));
    }
}

template<typename T>
static T parse(std::span<const uint8_t> buffer)
{
    T msg;
    if (!msg.ParseFromArray(buffer.data(), (int)buffer.size())) {
        throw std::runtime_error("Could not parse state request");
    }
    return msg;
}

std::string StateServer::recvSize(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateRequest>(buffer);
    SPDLOG_TRACE("Size {}/{}", request.user(), request.key());
    KV_FROM_REQUEST(request)
    faabric::StateSizeResponse response;
    response.set_user(kv->user);
    response.set_key(kv->key);
    response.set_statesize(kv->size());
    return response.SerializeAsString();
}

std::string StateServer::recvPull(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateChunkRequest>(buffer);
    SPDLOG_TRACE("Pull {}/{} ({}->{})", request.user(), request.key(), request.offset(), request.offset() + request.chunksize());
    KV_FROM_REQUEST(request)
    uint64_t chunkOffset = request.offset();
    uint64_t chunkLen = request.chunksize();
    if (chunkOffset + chunkLen > kv->size()) {
        SPDLOG_ERROR("Pull request {}/{} out of range ({}+{} > {})", request.user(), request.key(), chunkOffset, chunkLen, kv->size());
        throw std::runtime_error("State pull out of range");
    }
    faabric::StatePart response;
    std::string* data = response.mutable_data();
    data->resize(chunkLen);
    kv->getChunk((long)chunkOffset, BYTES(data->data()), chunkLen);
    response.set_user(request.user());
    response.set_key(request.key());
    response.set_offset(chunkOffset);
    return response.SerializeAsString();
}

std::string StateServer::recvPush(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StatePart>(buffer);
    SPDLOG_TRACE("Push {}/{} ({}->{})", request.user(), request.key(), request.offset(), request.offset() + request.data().size());
    KV_FROM_REQUEST(request)
    kv->setChunk((long)request.offset(), BYTES_CONST(request.data().data()), request.data().size());
    return faabric::EmptyResponse().SerializeAsString();
}

std::string StateServer::recvAppend(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateRequest>(buffer);
    SPDLOG_TRACE("Append {}/{}", request.user(), request.key());
    KV_FROM_REQUEST(request)
    kv->append(BYTES_CONST(request.data().data()), request.data().size());
    return faabric::EmptyResponse().SerializeAsString();
}

std::string StateServer::recvPullAppended(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateAppendedRequest>(buffer);
    SPDLOG_TRACE("Pull appended {}/{}", request.user(), request.key());
    KV_FROM_REQUEST(request)
    faabric::StateAppendedResponse response;
    response.set_user(request.user());
    response.set_key(request.key());
    std::lock_guard<std::mutex> lk(kv->getAppendedMutex());
    auto& values = kv->getAppendedValues();
    if (request.nvalues() > values.size()) {
        throw std::runtime_error("Pulling more appended values than exist");
    }
    for (uint32_t i = 0; i < request.nvalues(); i++) {
        AppendedInMemoryState& v = values.at(i);
        response.add_values()->set_data(v.data.get(), v.length);
    }
    return response.SerializeAsString();
}

std::string StateServer::recvDelete(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateRequest>(buffer);
    SPDLOG_TRACE("Delete {}/{}", request.user(), request.key());
    state.deleteKV(request.user(), request.key());
    return faabric::EmptyResponse().SerializeAsString();
}

std::string StateServer::recvClearAppended(std::span<const uint8_t> buffer)
{
    auto request = parse<faabric::StateRequest>(buffer);
    SPDLOG_TRACE("Clear appended {}/{}", request.user(), request.key());
    KV_FROM_REQUEST(request)
    kv->clearAppended();
    return faabric::EmptyResponse().SerializeAsString();
}

} // namespace faabric::state
