// Page-level memory helpers (reference: include/faabric/util/memory.h:16-97,
// src/util/memory.cpp:15-256) plus device-memory regions.
#pragma once

#include <cstdint>
#include <functional>
#include <memory>
#include <span>
#include <string>
#include <unistd.h>
#include <vector>

namespace faabric::util {

// dst[i] |= src[i]
void mergeManyDirtyPages(std::vector<char>& dest,
                         const std::vector<std::vector<char>>& source);

void mergeDirtyPages(std::vector<char>& dest, const std::vector<char>& source);

// -------------------------
// Alignment
// -------------------------
struct AlignedChunk
{
    long originalOffset = 0;
    long originalLength = 0;
    long nBytesOffset = 0;
    long nBytesLength = 0;
    long nPagesOffset = 0;
    long nPagesLength = 0;
    long offsetRemainder = 0;
};

static const long HOST_PAGE_SIZE = sysconf(_SC_PAGESIZE);

bool isPageAligned(const void* ptr);

size_t getRequiredHostPages(size_t nBytes);

size_t getRequiredHostPagesRoundDown(size_t nBytes);

size_t alignOffsetDown(size_t offset);

AlignedChunk getPageAlignedChunk(long offset, long length);

// -------------------------
// Allocation
// -------------------------
typedef std::unique_ptr<uint8_t[], std::function<void(uint8_t*)>> MemoryRegion;

MemoryRegion allocatePrivateMemory(size_t size);

MemoryRegion allocateSharedMemory(size_t size);

// PROT_NONE reservation that can later be claimed page by page
MemoryRegion allocateVirtualMemory(size_t size);

void claimVirtualMemory(std::span<uint8_t> region);

void mapMemoryPrivate(std::span<uint8_t> target, int fd);

void mapMemoryShared(std::span<uint8_t> target, int fd);

void resizeFd(int fd, size_t size);

void writeToFd(int fd, off_t offset, std::span<const uint8_t> data);

int createFd(size_t size, const std::string& fdLabel);

void appendDataToFd(int fd, std::span<uint8_t> data);

// -------------------------
// Device memory (B200)
// -------------------------
// Owning handle of cudaMalloc'd (or pinned-host) memory; empty on CPU boxes.
struct DeviceRegion
{
    uint8_t* ptr = nullptr;
    size_t size = 0;
    int device = -1;
    bool pinnedHost = false;

    DeviceRegion() = default;
    DeviceRegion(const DeviceRegion&) = delete;
    DeviceRegion& operator=(const DeviceRegion&) = delete;
    DeviceRegion(DeviceRegion&& o) noexcept;
    DeviceRegion& operator=(DeviceRegion&& o) noexcept;
    ~DeviceRegion();

    bool valid() const { return ptr != nullptr; }
    void release();
};

// Throws std::runtime_error if no device / allocation failure
DeviceRegion allocateDeviceMemory(size_t size, int device);

DeviceRegion allocatePinnedHostMemory(size_t size);

} // namespace faabric::util
