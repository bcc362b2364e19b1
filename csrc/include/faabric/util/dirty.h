// Dirty-page tracking.  Host memory: four interchangeable trackers selected by
// DIRTY_TRACKING_MODE (none | segfault | softpte | uffd[-wp|-thread|-thread-wp]),
// same contract as the reference (include/faabric/util/dirty.h:24-236): every
// mode reports the same pages for the same writes.  Device memory has no page
// faults to hook, so DeviceCompareDirtyTracker diffs against the base image with
// an sm_100a kernel (csrc/kernels/snapshot_kernels.cu: dirtyScanKernel).
#pragma once

#include <cstdint>
#include <memory>
#include <span>
#include <string>
#include <vector>

namespace faabric::util {

// Per-page flags are chars (0/1) like the reference, so that they can be merged
// and shipped around as plain byte vectors.
class DirtyTracker
{
  public:
    virtual ~DirtyTracker() = default;

    virtual void clearAll() = 0;

    virtual std::string getType() = 0;

    virtual void startTracking(std::span<uint8_t> region) = 0;

    virtual void stopTracking(std::span<uint8_t> region) = 0;

    virtual std::vector<char> getDirtyPages(std::span<uint8_t> region) = 0;

    virtual void startThreadLocalTracking(std::span<uint8_t> region) = 0;

    virtual void stopThreadLocalTracking(std::span<uint8_t> region) = 0;

    virtual std::vector<char> getThreadLocalDirtyPages(
      std::span<uint8_t> region) = 0;

    virtual std::vector<char> getBothDirtyPages(std::span<uint8_t> region) = 0;
};

// Marks everything dirty (cheap to "track", expensive to diff)
class NoneDirtyTracker final : public DirtyTracker
{
  public:
    void clearAll() override;
    std::string getType() override { return "none"; }
    void startTracking(std::span<uint8_t> region) override;
    void stopTracking(std::span<uint8_t> region) override;
    std::vector<char> getDirtyPages(std::span<uint8_t> region) override;
    void startThreadLocalTracking(std::span<uint8_t> region) override;
    void stopThreadLocalTracking(std::span<uint8_t> region) override;
    std::vector<char> getThreadLocalDirtyPages(
      std::span<uint8_t> region) override;
    std::vector<char> getBothDirtyPages(std::span<uint8_t> region) override;

  private:
    std::vector<char> dirtyPages;
};

// mprotect(PROT_READ) + SIGSEGV handler: first write to a page faults, the
// handler flags it and re-enables writes.
class SegfaultDirtyTracker final : public DirtyTracker
{
  public:
    SegfaultDirtyTracker();
    void clearAll() override;
    std::string getType() override { return "segfault"; }
    void startTracking(std::span<uint8_t> region) override;
    void stopTracking(std::span<uint8_t> region) override;
    std::vector<char> getDirtyPages(std::span<uint8_t> region) override;
    void startThreadLocalTracking(std::span<uint8_t> region) override;
    void stopThreadLocalTracking(std::span<uint8_t> region) override;
    std::vector<char> getThreadLocalDirtyPages(
      std::span<uint8_t> region) override;
    std::vector<char> getBothDirtyPages(std::span<uint8_t> region) override;

    // SIGSEGV handler
    static void handler(int sig, void* info, void* context) noexcept;

  private:
    void setUpSignalHandler();
};

// /proc/self/clear_refs + pagemap soft-dirty bit (bit 55)
class SoftPTEDirtyTracker final : public DirtyTracker
{
  public:
    SoftPTEDirtyTracker();
    ~SoftPTEDirtyTracker() override;
    void clearAll() override;
    std::string getType() override { return "softpte"; }
    void startTracking(std::span<uint8_t> region) override;
    void stopTracking(std::span<uint8_t> region) override;
    std::vector<char> getDirtyPages(std::span<uint8_t> region) override;
    void startThreadLocalTracking(std::span<uint8_t> region) override;
    void stopThreadLocalTracking(std::span<uint8_t> region) override;
    std::vector<char> getThreadLocalDirtyPages(
      std::span<uint8_t> region) override;
    std::vector<char> getBothDirtyPages(std::span<uint8_t> region) override;

    // True if the running kernel exposes soft-dirty bits
    static bool isSupported();

  private:
    int clearRefsFd = -1;
    int pagemapFd = -1;
};

// userfaultfd write-protect tracking, faults drained by an event thread
class UffdDirtyTracker final : public DirtyTracker
{
  public:
    explicit UffdDirtyTracker(const std::string& modeIn);
    ~UffdDirtyTracker() override;
    void clearAll() override;
    std::string getType() override { return mode; }
    void startTracking(std::span<uint8_t> region) override;
    void stopTracking(std::span<uint8_t> region) override;
    std::vector<char> getDirtyPages(std::span<uint8_t> region) override;
    void startThreadLocalTracking(std::span<uint8_t> region) override;
    void stopThreadLocalTracking(std::span<uint8_t> region) override;
    std::vector<char> getThreadLocalDirtyPages(
      std::span<uint8_t> region) override;
    std::vector<char> getBothDirtyPages(std::span<uint8_t> region) override;

    static bool isSupported();

  private:
    std::string mode;
    struct Impl;
    std::unique_ptr<Impl> impl;
};

// GPU memory: compare against a base image on the device
class DeviceCompareDirtyTracker
{
  public:
    // Returns one char per 4 KiB page of [mem, mem+size) that differs from base.
    // Both pointers are device pointers on `device`; synchronises `stream`.
    static std::vector<char> getDirtyPages(const uint8_t* mem,
                                           const uint8_t* base,
                                           size_t size,
                                           int device,
                                           void* stream = nullptr);

    // Same but leaves the flags on the device (uint8 per page); async
    static void getDirtyPagesDevice(const uint8_t* mem,
                                    const uint8_t* base,
                                    size_t size,
                                    uint8_t* pageFlagsDev,
                                    uint64_t* countDev,
                                    void* stream);
};

std::shared_ptr<DirtyTracker> getDirtyTracker();

// Re-reads the mode from the config (tests switch modes)
void resetDirtyTracker();

} // namespace faabric::util
