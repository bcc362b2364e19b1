// Snapshots: registry, RPCs and device-resident images.
//
// One header per module: the per-class headers of the reference's layout
// (faabric/snapshot/*.h) forward here, so either include style works.
#pragma once

#include <faabric/device/comm_abi.h>
#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/util/memory.h>
#include <faabric/util/snapshot.h>

#include <cstdint>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

// ==========================================================================
// snapshot/DeviceSnapshot.h
// ==========================================================================
// A snapshot whose image lives in GPU memory (HBM).  The hot path of the
// reference's fork-join threading — diff the executor memory against the
// image, apply the merge regions and push to the main copy
// (src/executor/Executor.cpp:684-730 + src/snapshot/SnapshotClient.cpp:76-171)
// — is ONE kernel here (snapshotDiffPushKernel) that writes straight into the
// main image, which may be peer-mapped memory of another GPU.



namespace faabric::snapshot {

struct DeviceDiffStats
{
    uint64_t diffBytes = 0;
    uint64_t pagesWithDiffs = 0;
};

// What travels over the control plane INSTEAD of the image when both ends are
// GPU hosts of one box: where the image lives and how to map it.  The
// reference serialises the whole image into the request
// (src/snapshot/SnapshotClient.cpp:76-108).
struct DeviceSnapshotDescriptor
{
    uint64_t size = 0;
    int device = -1;
    int ownerPid = 0;
    uint64_t devicePtr = 0;  // valid inside the owner process
    std::string ipcHandle;   // cudaIpcMemHandle_t bytes (empty: not exportable)
    std::vector<faabric::util::SnapshotMergeRegion> mergeRegions;
};

class DeviceSnapshot
{
  public:
    // Allocates `size` bytes on `device` (zeroed)
    DeviceSnapshot(size_t sizeIn, int deviceIn);

    // Wraps memory owned elsewhere (e.g. a symmetric-heap allocation that is
    // peer-mapped on the other GPUs)
    DeviceSnapshot(uint8_t* devicePtr, size_t sizeIn, int deviceIn);

    ~DeviceSnapshot();

    size_t getSize() const { return size; }

    int getDevice() const { return device; }

    uint8_t* getDevicePtr() const { return image; }

    // Host <-> device copies (synchronous)
    void copyInData(std::span<const uint8_t> hostData, uint64_t offset = 0);

    std::vector<uint8_t> getDataCopy(uint64_t offset, size_t n);

    std::vector<uint8_t> getDataCopy() { return getDataCopy(0, size); }

    // Restore: device-to-device copy of the image into executor memory
    void restoreTo(uint8_t* deviceMem, size_t n, void* stream = nullptr);

    // ---- spill / reload (checkpoint of a device image) ----
    // Host snapshot with the same bytes and merge regions (D2H copy)
    std::shared_ptr<faabric::util::SnapshotData> spillToHost();

    // New device image on `device` from a host snapshot (H2D copy)
    static std::shared_ptr<DeviceSnapshot> fromHost(faabric::util::SnapshotData& host, int device);

    void writeToFile(const std::string& path) { spillToHost()->writeToFile(path); }

    static std::shared_ptr<DeviceSnapshot> readFromFile(const std::string& path, int device)
    {
        return fromHost(*faabric::util::SnapshotData::readFromFile(path), device);
    }

    // ---- merge regions (same semantics as SnapshotData) ----
    void addMergeRegion(uint64_t offset,
                        size_t length,
                        faabric::util::SnapshotDataType dataType,
                        faabric::util::SnapshotMergeOperation operation);

    void clearMergeRegions();

    std::vector<faabric::util::SnapshotMergeRegion> getMergeRegions();

    // Fused diff + merge + push.  `mem` is the executor's (updated) memory on
    // this snapshot's device, this snapshot is its base image; results land in
    // `mainImage` (local or peer-mapped).  dirtyPagesDev: optional device
    // uint8[nPages].  If updateBase, this image is advanced to `mem` too.
    // Asynchronous on `stream`; stats are read back by getLastStats().
    void diffAndPush(const uint8_t* mem,
                     size_t memSize,
                     uint8_t* mainImage,
                     const uint8_t* dirtyPagesDev = nullptr,
                     bool updateBase = false,
                     void* stream = nullptr);

    // Synchronises `stream` and returns the counters of the last diffAndPush
    DeviceDiffStats getLastStats(void* stream = nullptr);

    // Applies host-side diffs (e.g. received over the control plane)
    void applyDiffs(const std::vector<faabric::util::SnapshotDiff>& diffs,
                    void* stream = nullptr);

    // GPU dirty-page detection of `mem` against this image
    std::vector<char> getDirtyPages(const uint8_t* mem, size_t memSize);

    // ---- delta codec on device images (util/delta.h command stream) ----
    // Delta that turns this image into `mem` (device memory on this image's
    // GPU, memSize <= size): the page compare and the XOR run on the GPU, only
    // the changed pages cross PCIe.  Byte-identical to
    // faabric::util::serializeDelta on host copies of the two buffers.
    // Settings without 4 KiB pages go through host copies.
    std::vector<uint8_t> serializeDelta(const faabric::util::DeltaSettings& cfg, const uint8_t* mem, size_t memSize);

    // Applies a delta to this image in place (XOR / overwrite runs become one
    // batch of device diffs)
    void applyDelta(const std::vector<uint8_t>& delta, void* stream = nullptr);

    // ---- control-plane descriptor (cross-process mapping through CUDA IPC) ----
    DeviceSnapshotDescriptor describe();

    // Maps an image owned by ANOTHER process of this box (throws if the
    // descriptor carries no IPC handle)
    static std::shared_ptr<DeviceSnapshot> fromDescriptor(const DeviceSnapshotDescriptor& desc);

    // ---- incremental synchronisation (THREADS fork-join after the first) ----
    // Every 4 KiB page of an image this process owns carries a stamp: the last
    // fork (2k: the main thread's refresh) or join (2k+1: a host's merge) that
    // changed it.  A copy that was identical to the image at stamp s is brought
    // up to date by pulling the pages stamped later than s.
    // nullptr when the image has no stamps (wrapped or mapped from elsewhere)
    uint32_t* pageStamps();

    // Starts fork k: returns 2k.  currentForkStamp() repeats it.
    uint32_t beginFork();

    uint32_t currentForkStamp() const { return 2 * forkCounter.load(); }

    // Copies that synchronised before this stamp must be copied whole again
    // (the image was rewritten by something that does not stamp pages)
    uint32_t fullMutationStamp() const { return fullMutation.load(); }

    // Distinguishes images that reuse an address
    uint64_t uid() const { return uniqueId; }

    // The image was (or may have been) changed without stamping pages
    void markRewritten() { noteFullMutation(); }

    // image := mem wherever they differ, stamping the changed pages with
    // `stamp` (asynchronous on `stream`, which must belong to this image's GPU)
    void syncPagesFrom(const uint8_t* mem, size_t n, uint32_t stamp, void* stream = nullptr);

    // Pages stamped later than `since` are copied into dst1 (and dst2); runs on
    // `onDevice`, the GPU that owns the destinations.  Asynchronous.
    void pullChangedPages(uint8_t* dst1, uint8_t* dst2, uint32_t since, size_t n, int onDevice, void* stream = nullptr);

    // Pages moved by the launches above since the counter was last read
    // (synchronises `stream`)
    uint64_t takePageCopyCount(int onDevice, void* stream = nullptr);

    // diffAndPush into an image that keeps stamps: pages this launch changes
    // there are stamped with `stamp`
    void setPushStamps(uint32_t* stampsOfTarget, uint32_t stamp)
    {
        pushStamps = stampsOfTarget;
        pushStamp = stamp;
    }

    // How many fused diff+push launches this image has issued (tests, metrics)
    uint64_t getDiffPushCount() const { return diffPushCount; }

    static uint64_t getGlobalDiffPushCount();

  private:
    size_t size = 0;
    int device = 0;
    uint8_t* image = nullptr;
    faabric::util::DeviceRegion owned;

    std::mutex mx;
    std::vector<faabric::util::SnapshotMergeRegion> mergeRegions;

    // Uploaded, gap-filled regions (rebuilt lazily when regions change)
    bool regionsDirty = true;
    faabric::util::DeviceRegion regionsDev;
    faabric::util::DeviceRegion typedIdxDev;
    faabric::util::DeviceRegion statsDev;
    int nRegionsDev = 0;
    int nTypedDev = 0;

    void uploadRegions();

    void* ipcMapped = nullptr;
    uint64_t diffPushCount = 0;

    faabric::util::DeviceRegion stampsDev;
    std::atomic<uint32_t> forkCounter{ 0 };
    std::atomic<uint32_t> fullMutation{ 0 };
    uint64_t uniqueId = 0;
    uint32_t* pushStamps = nullptr;
    uint32_t pushStamp = 0;
    // per-GPU counters for the page kernels (they run where the copies live)
    std::map<int, faabric::util::DeviceRegion> pageStatsDev;

    uint64_t* pageStatsOn(int onDevice);

    void noteFullMutation() { fullMutation.store(2 * forkCounter.load() + 1); }
};

}

// ==========================================================================
// snapshot/SnapshotApi.h
// ==========================================================================
namespace faabric::snapshot {
enum SnapshotCalls
{
    NoSnapshotCall = 0,
    PushSnapshot = 1,
    PushSnapshotUpdate = 2,
    DeleteSnapshot = 3,
    ThreadResult = 4,
};
}

// ==========================================================================
// snapshot/SnapshotClient.h
// ==========================================================================
namespace faabric::snapshot {

// -----------------------------------
// Mocking (reference: src/snapshot/SnapshotClient.cpp:18-64)
// -----------------------------------
struct MockSnapshotUpdate
{
    std::vector<faabric::util::SnapshotDiff> diffs;
    // diffs above are non-owning: the payloads are kept here
    std::vector<std::vector<uint8_t>> diffData;
    std::vector<faabric::util::SnapshotMergeRegion> mergeRegions;
};

std::vector<
  std::pair<std::string, std::shared_ptr<faabric::util::SnapshotData>>>
getSnapshotPushes();

std::vector<std::pair<std::string, std::shared_ptr<MockSnapshotUpdate>>>
getSnapshotDiffPushes();

std::vector<std::pair<std::string, std::string>> getSnapshotDeletes();

std::vector<std::pair<std::string, std::tuple<int, int, std::string, int>>>
getThreadResults();

// (host, key, descriptor) of every device-image descriptor pushed in mock mode
std::vector<std::tuple<std::string, std::string, DeviceSnapshotDescriptor>>
getDeviceSnapshotPushes();

void clearMockSnapshotRequests();

// -----------------------------------
// Client
// -----------------------------------
class SnapshotClient final : public faabric::transport::MessageEndpointClient
{
  public:
    explicit SnapshotClient(const std::string& hostIn);

    void pushSnapshot(const std::string& key,
                      std::shared_ptr<faabric::util::SnapshotData> data);

    void pushSnapshotUpdate(
      std::string snapshotKey,
      const std::shared_ptr<faabric::util::SnapshotData>& data,
      const std::vector<faabric::util::SnapshotDiff>& diffs);

    void deleteSnapshot(const std::string& key);

    void pushThreadResult(uint32_t appId,
                          uint32_t messageId,
                          int returnValue,
                          const std::string& key,
                          const std::vector<faabric::util::SnapshotDiff>& diffs);

    // ---- GPU hosts: control descriptors only, the bytes stay in HBM ----
    // "Push" of a device-resident image: the receiver learns where it lives
    void pushDeviceSnapshot(const std::string& key,
                            const DeviceSnapshotDescriptor& desc);

    // Thread result whose diffs were already merged into the main image by
    // the fused kernel on the sender's GPU
    void pushDeviceThreadResult(uint32_t appId,
                                uint32_t messageId,
                                int returnValue,
                                const std::string& key,
                                uint64_t diffBytes);

  private:
    // True when `host` is served by this process and its registry (ours)
    // already holds this very object under `key`
    bool receiverSharesRegistry(const std::string& key,
                                const std::shared_ptr<faabric::util::SnapshotData>& data);
};

std::shared_ptr<SnapshotClient> getSnapshotClient(const std::string& host);

void clearSnapshotClients();

}

// ==========================================================================
// snapshot/SnapshotRegistry.h
// ==========================================================================
namespace faabric::snapshot {

class DeviceSnapshot;

// key -> snapshot (host images and device images live side by side)
class SnapshotRegistry
{
  public:
    SnapshotRegistry() = default;

    std::shared_ptr<faabric::util::SnapshotData> getSnapshot(
      const std::string& key);

    bool snapshotExists(const std::string& key);

    void registerSnapshot(const std::string& key,
                          std::shared_ptr<faabric::util::SnapshotData> data);

    void deleteSnapshot(const std::string& key);

    size_t getSnapshotCount();

    // ---- device-resident images ----
    std::shared_ptr<DeviceSnapshot> getDeviceSnapshot(const std::string& key);

    bool deviceSnapshotExists(const std::string& key);

    void registerDeviceSnapshot(const std::string& key,
                                std::shared_ptr<DeviceSnapshot> data);

    void deleteDeviceSnapshot(const std::string& key);

    // Descriptors of device images owned elsewhere (resolved to a mapped
    // DeviceSnapshot by getDeviceSnapshot on first use)
    void registerDeviceDescriptor(const std::string& key,
                                  const DeviceSnapshotDescriptor& desc);

    bool deviceDescriptorExists(const std::string& key);

    DeviceSnapshotDescriptor getDeviceDescriptor(const std::string& key);

    void clear();

    // ---- on-disk checkpoints: one file per snapshot under `dir` (created if
    // missing), named by the hex-encoded key; device images are spilled
    // through the host.  Returns the number of files written / snapshots
    // registered.  `device` < 0 restores everything as host snapshots,
    // otherwise images saved from device memory return to that GPU. ----
    size_t checkpointToDir(const std::string& dir);

    size_t restoreFromDir(const std::string& dir, int device = -1);

  private:
    std::shared_mutex snapshotsMx;
    std::unordered_map<std::string, std::shared_ptr<faabric::util::SnapshotData>>
      snapshotMap;
    std::unordered_map<std::string, std::shared_ptr<DeviceSnapshot>> deviceMap;
    std::unordered_map<std::string, DeviceSnapshotDescriptor> descriptorMap;
};

SnapshotRegistry& getSnapshotRegistry();

}

// ==========================================================================
// snapshot/SnapshotServer.h
// ==========================================================================
namespace faabric::snapshot {

class SnapshotServer final : public faabric::transport::MessageEndpointServer
{
  public:
    SnapshotServer();

  protected:
    void doAsyncRecv(transport::Message& message) override;

    std::string doSyncRecv(transport::Message& message) override;

    std::string recvPushSnapshot(std::span<const uint8_t> buffer);

    std::string recvPushSnapshotUpdate(std::span<const uint8_t> buffer);

    std::string recvThreadResult(transport::Message& message);

    void recvDeleteSnapshot(std::span<const uint8_t> buffer);

  private:
    faabric::snapshot::SnapshotRegistry& reg;
};

}

