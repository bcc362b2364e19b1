// Device-resident state values (see state/DeviceStateKeyValue.h)
#include <faabric/state/State.h>
#include <faabric/util/logging.h>

#include "launch_api.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>

namespace faabric::state {

namespace {
struct DevGuard
{
    int prev = -1;
    explicit DevGuard(int dev)
    {
        cudaGetDevice(&prev);
        cudaSetDevice(dev);
    }
    ~DevGuard()
    {
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
};

void ck(cudaError_t e, const char* what)
{
    if (e != cudaSuccess) {
        throw StateKeyValueException(std::string(what) + ": " + cudaGetErrorString(e));
    }
}
}

DeviceStateKeyValue::DeviceStateKeyValue(std::string userIn,
                                         std::string keyIn,
                                         size_t sizeIn,
                                         int deviceIn,
                                         std::shared_ptr<DeviceStateKeyValue> mainIn)
  : user(std::move(userIn))
  , key(std::move(keyIn))
  , valueSize(sizeIn)
  , device(deviceIn)
  , main(std::move(mainIn))
{
    if (valueSize == 0) {
        throw StateKeyValueException("Device state " + user + "/" + key + " has no size");
    }
    if (main != nullptr && main->size() != valueSize) {
        throw StateKeyValueException("Device state " + user + "/" + key + " size mismatch with its main copy");
    }
    const size_t nBlocks = (valueSize + FB_STATE_BLOCK_BYTES - 1) / FB_STATE_BLOCK_BYTES;
    data = faabric::util::allocateDeviceMemory(valueSize, device);
    mask = faabric::util::allocateDeviceMemory(nBlocks, device);
    stats = faabric::util::allocateDeviceMemory(64, device);
    DevGuard g(device);
    ck(cudaMemset(data.ptr, 0, valueSize), "state memset");
    ck(cudaMemset(mask.ptr, 0, nBlocks), "state mask memset");
    const size_t nChunks = (valueSize + STATE_STREAMING_CHUNK_SIZE - 1) / STATE_STREAMING_CHUNK_SIZE;
    // the main copy is by definition up to date
    pulledChunks.assign(nChunks, main == nullptr ? 1 : 0);
    if (main != nullptr && main->getDevice() != device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(main->getDevice(), 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
            SPDLOG_WARN("No peer access from GPU {} to GPU {}: {}", device, main->getDevice(), cudaGetErrorString(e));
        }
        cudaGetLastError();
    }
}

DeviceStateKeyValue::~DeviceStateKeyValue() = default;

void DeviceStateKeyValue::checkRange(long offset, size_t length) const
{
    if (offset < 0 || (size_t)offset + length > valueSize) {
        throw StateKeyValueException("Device state chunk out of bounds: " + std::to_string(offset) + " + " +
                                     std::to_string(length) + " > " + std::to_string(valueSize));
    }
}

void DeviceStateKeyValue::pull(void* stream)
{
    if (main == nullptr) {
        return;
    }
    std::lock_guard<std::mutex> lk(mx);
    DevGuard g(device);
    ck(cudaMemcpyAsync(data.ptr, main->getDevicePtr(), valueSize, cudaMemcpyDefault, (cudaStream_t)stream), "state pull");
    ck(cudaStreamSynchronize((cudaStream_t)stream), "state pull sync");
    std::fill(pulledChunks.begin(), pulledChunks.end(), 1);
    bytesPulled += valueSize;
}

bool DeviceStateKeyValue::isChunkPulled(long offset, size_t length)
{
    checkRange(offset, length);
    if (length == 0) {
        return true;
    }
    std::lock_guard<std::mutex> lk(mx);
    size_t c0 = (size_t)offset / STATE_STREAMING_CHUNK_SIZE;
    size_t c1 = ((size_t)offset + length - 1) / STATE_STREAMING_CHUNK_SIZE;
    for (size_t c = c0; c <= c1; c++) {
        if (!pulledChunks[c]) {
            return false;
        }
    }
    return true;
}

void DeviceStateKeyValue::pullChunk(long offset, size_t length, void* stream)
{
    checkRange(offset, length);
    if (main == nullptr || length == 0) {
        return;
    }
    std::lock_guard<std::mutex> lk(mx);
    DevGuard g(device);
    size_t c0 = (size_t)offset / STATE_STREAMING_CHUNK_SIZE;
    size_t c1 = ((size_t)offset + length - 1) / STATE_STREAMING_CHUNK_SIZE;
    // coalesce neighbouring missing chunks into one peer copy each
    size_t c = c0;
    bool any = false;
    while (c <= c1) {
        if (pulledChunks[c]) {
            c++;
            continue;
        }
        size_t e = c;
        while (e <= c1 && !pulledChunks[e]) {
            pulledChunks[e] = 1;
            e++;
        }
        size_t b0 = c * STATE_STREAMING_CHUNK_SIZE;
        size_t b1 = std::min(valueSize, e * (size_t)STATE_STREAMING_CHUNK_SIZE);
        ck(cudaMemcpyAsync(data.ptr + b0, main->getDevicePtr() + b0, b1 - b0, cudaMemcpyDefault, (cudaStream_t)stream),
           "state chunk pull");
        bytesPulled += b1 - b0;
        any = true;
        c = e;
    }
    if (any) {
        ck(cudaStreamSynchronize((cudaStream_t)stream), "state chunk pull sync");
    }
}

void DeviceStateKeyValue::flagChunkDirty(long offset, long len, void* stream)
{
    checkRange(offset, (size_t)len);
    if (len <= 0) {
        return;
    }
    DevGuard g(device);
    uint64_t b0 = (uint64_t)offset / FB_STATE_BLOCK_BYTES;
    uint64_t b1 = ((uint64_t)offset + (uint64_t)len - 1) / FB_STATE_BLOCK_BYTES;
    ck(fb::launchStateFlagRange(mask.ptr, b0, b1 - b0 + 1, (cudaStream_t)stream), "state flag");
}

void DeviceStateKeyValue::flagDirty(void* stream)
{
    flagChunkDirty(0, (long)valueSize, stream);
}

uint64_t DeviceStateKeyValue::pushPartial(void* stream)
{
    if (main == nullptr) {
        // the main copy IS the value: just forget the flags
        DevGuard g(device);
        ck(cudaMemsetAsync(mask.ptr, 0, mask.size, (cudaStream_t)stream), "state mask clear");
        return 0;
    }
    std::lock_guard<std::mutex> lk(mx);
    DevGuard g(device);
    auto s = (cudaStream_t)stream;
    ck(cudaMemsetAsync(stats.ptr, 0, 16, s), "state stats");
    ck(fb::launchStatePushDirty(mask.ptr, data.ptr, main->getDevicePtr(), valueSize, (uint64_t*)stats.ptr, 0, s),
       "state push kernel");
    pushLaunches++;
    uint64_t blocks = 0;
    ck(cudaMemcpyAsync(&blocks, stats.ptr, 8, cudaMemcpyDeviceToHost, s), "state stats read");
    ck(cudaStreamSynchronize(s), "state push sync");
    return std::min<uint64_t>(blocks * FB_STATE_BLOCK_BYTES, valueSize);
}

void DeviceStateKeyValue::pushFull(void* stream)
{
    if (main == nullptr) {
        return;
    }
    std::lock_guard<std::mutex> lk(mx);
    DevGuard g(device);
    ck(cudaMemcpyAsync(main->getDevicePtr(), data.ptr, valueSize, cudaMemcpyDefault, (cudaStream_t)stream), "state push full");
    ck(cudaMemsetAsync(mask.ptr, 0, mask.size, (cudaStream_t)stream), "state mask clear");
    ck(cudaStreamSynchronize((cudaStream_t)stream), "state push full sync");
}

std::vector<DeviceStateRun> DeviceStateKeyValue::getDirtyChunks(void* stream)
{
    std::lock_guard<std::mutex> lk(mx);
    DevGuard g(device);
    auto s = (cudaStream_t)stream;
    const uint64_t nBlocks = (valueSize + FB_STATE_BLOCK_BYTES - 1) / FB_STATE_BLOCK_BYTES;
    // worst case: every other block dirty
    const uint32_t maxOut = (uint32_t)std::min<uint64_t>(nBlocks / 2 + 1, 1u << 22);
    auto out = faabric::util::allocateDeviceMemory((size_t)maxOut * sizeof(FbDiffDesc), device);
    auto cnt = faabric::util::allocateDeviceMemory(16, device);
    ck(cudaMemsetAsync(cnt.ptr, 0, 16, s), "state runs count");
    ck(fb::launchChunkRuns(mask.ptr, nBlocks, FB_STATE_BLOCK_BYTES, valueSize, (FbDiffDesc*)out.ptr, maxOut, (uint32_t*)cnt.ptr, s),
       "state runs kernel");
    uint32_t n = 0;
    ck(cudaMemcpyAsync(&n, cnt.ptr, 4, cudaMemcpyDeviceToHost, s), "state runs count read");
    ck(cudaStreamSynchronize(s), "state runs sync");
    n = std::min(n, maxOut);
    std::vector<FbDiffDesc> descs(n);
    if (n > 0) {
        ck(cudaMemcpy(descs.data(), out.ptr, (size_t)n * sizeof(FbDiffDesc), cudaMemcpyDeviceToHost), "state runs read");
    }
    std::vector<DeviceStateRun> runs;
    runs.reserve(n);
    for (const auto& d : descs) {
        runs.push_back({ d.offset, d.length });
    }
    std::sort(runs.begin(), runs.end(), [](const DeviceStateRun& a, const DeviceStateRun& b) { return a.offset < b.offset; });
    return runs;
}

uint8_t* DeviceStateKeyValue::syncHostMirror(void* stream)
{
    std::lock_guard<std::mutex> lk(mx);
    if (!hostMirror.valid()) {
        hostMirror = faabric::util::allocatePinnedHostMemory(valueSize);
    }
    DevGuard g(device);
    ck(cudaMemcpyAsync(hostMirror.ptr, data.ptr, valueSize, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "state mirror");
    ck(cudaStreamSynchronize((cudaStream_t)stream), "state mirror sync");
    return hostMirror.ptr;
}

void DeviceStateKeyValue::get(uint8_t* buffer)
{
    getChunk(0, buffer, valueSize);
}

void DeviceStateKeyValue::getChunk(long offset, uint8_t* buffer, size_t length)
{
    checkRange(offset, length);
    pullChunk(offset, length);
    DevGuard g(device);
    ck(cudaMemcpy(buffer, data.ptr + offset, length, cudaMemcpyDeviceToHost), "state get");
}

void DeviceStateKeyValue::set(const uint8_t* buffer)
{
    setChunk(0, buffer, valueSize);
}

void DeviceStateKeyValue::setChunk(long offset, const uint8_t* buffer, size_t length)
{
    checkRange(offset, length);
    {
        DevGuard g(device);
        ck(cudaMemcpy(data.ptr + offset, buffer, length, cudaMemcpyHostToDevice), "state set");
    }
    // whole chunks written by the host need no pull any more
    {
        std::lock_guard<std::mutex> lk(mx);
        size_t first = ((size_t)offset + STATE_STREAMING_CHUNK_SIZE - 1) / STATE_STREAMING_CHUNK_SIZE;
        size_t last = ((size_t)offset + length) / STATE_STREAMING_CHUNK_SIZE;
        for (size_t c = first; c < last && c < pulledChunks.size(); c++) {
            pulledChunks[c] = 1;
        }
        if ((size_t)offset + length == valueSize && !pulledChunks.empty() &&
            (size_t)offset <= (pulledChunks.size() - 1) * (size_t)STATE_STREAMING_CHUNK_SIZE) {
            pulledChunks.back() = 1;
        }
    }
    flagChunkDirty(offset, (long)length);
    DevGuard g(device);
    ck(cudaStreamSynchronize(nullptr), "state set sync");
}

// ---------------------------------------------------------------------------
// registry
// ---------------------------------------------------------------------------
std::shared_ptr<DeviceStateKeyValue> State::getDeviceKV(const std::string& user,
                                                       const std::string& key,
                                                       size_t size,
                                                       int device)
{
    if (user.empty() || key.empty()) {
        throw StateKeyValueException("Attempting to access device state with empty user or key");
    }
    const std::string full = user + "_" + key;
    std::unique_lock<std::shared_mutex> lock(mapMutex);
    auto& perDevice = deviceKvMap[full];
    auto it = perDevice.find(device);
    if (it != perDevice.end()) {
        return it->second;
    }
    std::shared_ptr<DeviceStateKeyValue> mainCopy;
    auto mit = deviceKvMain.find(full);
    if (mit != deviceKvMain.end()) {
        mainCopy = perDevice.at(mit->second);
        if (size == 0) {
            size = mainCopy->size();
        }
    } else {
        deviceKvMain[full] = device;
        SPDLOG_DEBUG("GPU {} is the main copy of device state {}", device, full);
    }
    auto kv = std::make_shared<DeviceStateKeyValue>(user, key, size, device, mainCopy);
    perDevice[device] = kv;
    return kv;
}

void State::deleteDeviceKV(const std::string& user, const std::string& key)
{
    std::unique_lock<std::shared_mutex> lock(mapMutex);
    const std::string full = user + "_" + key;
    deviceKvMap.erase(full);
    deviceKvMain.erase(full);
}

size_t State::getDeviceKVCount()
{
    std::shared_lock<std::shared_mutex> lock(mapMutex);
    return deviceKvMap.size();
}

} // namespace faabric::state
