#include <faabric/util/logging.h>
#include <faabric/util/memory.h>

#include <cuda_runtime.h>

#include <cstring>
#include <fcntl.h>
#include <stdexcept>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace faabric::util {

void mergeManyDirtyPages(std::vector<char>& dest,
                         const std::vector<std::vector<char>>& source)
{
    for (const auto& s : source) {
        mergeDirtyPages(dest, s);
    }
}

void mergeDirtyPages(std::vector<char>& dest, const std::vector<char>& source)
{
    if (dest.size() < source.size()) {
        dest.resize(source.size(), 0);
    }
    for (size_t i = 0; i < source.size(); i++) {
        dest[i] |= source[i];
    }
}

bool isPageAligned(const void* ptr)
{
    return (((uintptr_t)ptr) & (HOST_PAGE_SIZE - 1)) == 0;
}

size_t getRequiredHostPages(size_t nBytes)
{
    return (nBytes + HOST_PAGE_SIZE - 1) / HOST_PAGE_SIZE;
}

size_t getRequiredHostPagesRoundDown(size_t nBytes)
{
    return nBytes / HOST_PAGE_SIZE;
}

size_t alignOffsetDown(size_t offset)
{
    return offset - (offset % HOST_PAGE_SIZE);
}

AlignedChunk getPageAlignedChunk(long offset, long length)
{
    AlignedChunk c;
    c.originalOffset = offset;
    c.originalLength = length;
    c.nPagesOffset = offset / HOST_PAGE_SIZE;
    c.nBytesOffset = c.nPagesOffset * HOST_PAGE_SIZE;
    c.offsetRemainder = offset - c.nBytesOffset;
    long adjusted = length + c.offsetRemainder;
    c.nPagesLength = (adjusted + HOST_PAGE_SIZE - 1) / HOST_PAGE_SIZE;
    c.nBytesLength = c.nPagesLength * HOST_PAGE_SIZE;
    return c;
}

static MemoryRegion doAlloc(size_t size, int prot, int flags)
{
    void* p = ::mmap(nullptr, size, prot, flags, -1, 0);
    if (p == MAP_FAILED) {
        SPDLOG_ERROR("mmap of {} bytes failed: {}", size, strerror(errno));
        throw std::runtime_error("Allocating memory with mmap failed");
    }
    auto deleter = [size](uint8_t* u) { ::munmap(u, size); };
    return MemoryRegion((uint8_t*)p, deleter);
}

MemoryRegion allocatePrivateMemory(size_t size)
{
    return doAlloc(size, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS);
}

MemoryRegion allocateSharedMemory(size_t size)
{
    return doAlloc(size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS);
}

MemoryRegion allocateVirtualMemory(size_t size)
{
    return doAlloc(size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS);
}

void claimVirtualMemory(std::span<uint8_t> region)
{
    if (::mprotect(region.data(), region.size(), PROT_READ | PROT_WRITE) != 0) {
        SPDLOG_ERROR("Failed claiming virtual memory: {}", strerror(errno));
        throw std::runtime_error("Failed claiming virtual memory");
    }
}

static void mapMemory(std::span<uint8_t> target, int fd, int flags)
{
    if (!isPageAligned(target.data())) {
        SPDLOG_ERROR("Mapping memory to non page-aligned address");
        throw std::runtime_error("Mapping memory to non page-aligned address");
    }
    if (fd <= 0) {
        SPDLOG_ERROR("Mapping invalid or zero fd ({})", fd);
        throw std::runtime_error("Invalid fd for mapping");
    }
    void* res = ::mmap(target.data(),
                       target.size(),
                       PROT_READ | PROT_WRITE,
                       flags | MAP_FIXED,
                       fd,
                       0);
    if (res == MAP_FAILED) {
        SPDLOG_ERROR("mmapping memory failed: {} ({})", errno, strerror(errno));
        throw std::runtime_error("mmapping memory failed");
    }
}

void mapMemoryPrivate(std::span<uint8_t> target, int fd)
{
    mapMemory(target, fd, MAP_PRIVATE);
}

void mapMemoryShared(std::span<uint8_t> target, int fd)
{
    mapMemory(target, fd, MAP_SHARED);
}

void resizeFd(int fd, size_t size)
{
    if (::ftruncate(fd, (off_t)size) != 0) {
        SPDLOG_ERROR("ftruncate failed with fd {}: {}", fd, strerror(errno));
        throw std::runtime_error("Failed to resize fd");
    }
}

void writeToFd(int fd, off_t offset, std::span<const uint8_t> data)
{
    size_t done = 0;
    while (done < data.size()) {
        ssize_t n =
          ::pwrite(fd, data.data() + done, data.size() - done, offset + done);
        if (n < 0) {
            if (errno == EINTR) {
                continue;
            }
            SPDLOG_ERROR("Write to fd {} failed: {}", fd, strerror(errno));
            throw std::runtime_error("Failed writing to fd");
        }
        done += (size_t)n;
    }
}

int createFd(size_t size, const std::string& fdLabel)
{
    int fd = ::memfd_create(fdLabel.c_str(), 0);
    if (fd == -1) {
        SPDLOG_ERROR("Failed to create memfd: {}", strerror(errno));
        throw std::runtime_error("Failed to create memfd");
    }
    if (size > 0) {
        resizeFd(fd, size);
    }
    return fd;
}

void appendDataToFd(int fd, std::span<uint8_t> data)
{
    off_t end = ::lseek(fd, 0, SEEK_END);
    if (end == -1) {
        throw std::runtime_error("lseek failed");
    }
    resizeFd(fd, (size_t)end + data.size());
    writeToFd(fd, end, data);
}

// ------------------------------------------------------------ device ---
DeviceRegion::DeviceRegion(DeviceRegion&& o) noexcept
{
    *this = std::move(o);
}

DeviceRegion& DeviceRegion::operator=(DeviceRegion&& o) noexcept
{
    if (this != &o) {
        release();
        ptr = o.ptr;
        size = o.size;
        device = o.device;
        pinnedHost = o.pinnedHost;
        o.ptr = nullptr;
        o.size = 0;
    }
    return *this;
}

DeviceRegion::~DeviceRegion()
{
    release();
}

void DeviceRegion::release()
{
    if (ptr == nullptr) {
        return;
    }
    if (pinnedHost) {
        cudaFreeHost(ptr);
    } else {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        cudaFree(ptr);
        if (prev >= 0) {
            cudaSetDevice(prev);
        }
    }
    cudaGetLastError();
    ptr = nullptr;
    size = 0;
}

DeviceRegion allocateDeviceMemory(size_t size, int device)
{
    DeviceRegion r;
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error("allocateDeviceMemory: no such CUDA device");
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, size);
    if (prev >= 0) {
        cudaSetDevice(prev);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error(std::string("cudaMalloc failed: ") +
                                 cudaGetErrorString(e));
    }
    r.ptr = (uint8_t*)p;
    r.size = size;
    r.device = device;
    return r;
}

DeviceRegion allocatePinnedHostMemory(size_t size)
{
    DeviceRegion r;
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, size, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error(std::string("cudaHostAlloc failed: ") +
                                 cudaGetErrorString(e));
    }
    r.ptr = (uint8_t*)p;
    r.size = size;
    r.pinnedHost = true;
    return r;
}

} // namespace faabric::util
