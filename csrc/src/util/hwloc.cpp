#include <faabric/util/config.h>
#include <faabric/util/environment.h>
#include <faabric/util/hwloc.h>
#include <faabric/util/logging.h>
#include <faabric/util/testing.h>

#include <cuda_runtime.h>

#include <fstream>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace faabric::util {

// Index of a set handed out when every CPU is taken and the process runs in
// test mode (reference: src/util/hwloc.cpp:77-82 does the same for small CI
// machines)
static constexpr int SHARED_CPU_IDX = -2;

static std::mutex cpuMx;
static std::vector<bool> cpuTaken;
static bool cpuInit = false;

static void initCpuSet()
{
    if (cpuInit) {
        return;
    }
    unsigned int n = getUsableCores();
    cpuTaken.assign(n, false);
    // OVERRIDE_FREE_CPU_START reserves the low CPUs for other tenants
    int start = getSystemConfig().overrideFreeCpuStart;
    for (int i = 0; i < start && i < (int)n; i++) {
        cpuTaken[i] = true;
    }
    cpuInit = true;
}

FaabricCpuSet::FaabricCpuSet(int cpuIdxIn)
  : cpuIdx(cpuIdxIn)
{
    CPU_ZERO(&cpuSet);
    if (cpuIdx >= 0) {
        CPU_SET(cpuIdx, &cpuSet);
    } else if (cpuIdx == SHARED_CPU_IDX) {
        // An overcommitted pin (test mode only) shares CPU 0 and owns nothing
        CPU_SET(0, &cpuSet);
    }
}

FaabricCpuSet::~FaabricCpuSet()
{
    if (cpuIdx < 0) {
        return;
    }
    std::lock_guard<std::mutex> lk(cpuMx);
    if (cpuIdx < (int)cpuTaken.size()) {
        cpuTaken[cpuIdx] = false;
    }
}

int getNumFreeCpus()
{
    std::lock_guard<std::mutex> lk(cpuMx);
    initCpuSet();
    int n = 0;
    for (bool t : cpuTaken) {
        n += t ? 0 : 1;
    }
    return n;
}

static std::unique_ptr<FaabricCpuSet> pinToOneOf(pthread_t thread,
                                                 const std::vector<int>& prefer)
{
    int chosen = -1;
    {
        std::lock_guard<std::mutex> lk(cpuMx);
        initCpuSet();
        for (int c : prefer) {
            if (c >= 0 && c < (int)cpuTaken.size() && !cpuTaken[c]) {
                chosen = c;
                break;
            }
        }
        if (chosen < 0) {
            for (size_t c = 0; c < cpuTaken.size(); c++) {
                if (!cpuTaken[c]) {
                    chosen = (int)c;
                    break;
                }
            }
        }
        if (chosen < 0 && isTestMode()) {
            chosen = SHARED_CPU_IDX;
        } else if (chosen < 0) {
            SPDLOG_ERROR("No free CPUs left to pin a thread to");
            throw std::runtime_error("No free CPUs to pin to");
        } else {
            cpuTaken[chosen] = true;
        }
    }
    auto set = std::make_unique<FaabricCpuSet>(chosen);
    int rc = pthread_setaffinity_np(thread, sizeof(cpu_set_t), set->get());
    if (rc != 0) {
        SPDLOG_WARN("pthread_setaffinity_np to CPU {} failed ({})", chosen, rc);
    }
    return set;
}

std::unique_ptr<FaabricCpuSet> pinThreadToFreeCpu(pthread_t thread)
{
    return pinToOneOf(thread, {});
}

// Parses "0-31,64-95" style cpulists
static std::vector<int> parseCpuList(const std::string& s)
{
    std::vector<int> out;
    size_t i = 0;
    while (i < s.size()) {
        size_t j = s.find(',', i);
        std::string part = s.substr(i, j == std::string::npos ? j : j - i);
        size_t dash = part.find('-');
        try {
            if (dash == std::string::npos) {
                out.push_back(std::stoi(part));
            } else {
                int a = std::stoi(part.substr(0, dash));
                int b = std::stoi(part.substr(dash + 1));
                for (int c = a; c <= b; c++) {
                    out.push_back(c);
                }
            }
        } catch (...) {
        }
        if (j == std::string::npos) {
            break;
        }
        i = j + 1;
    }
    return out;
}

std::unique_ptr<FaabricCpuSet> pinThreadNearGpu(pthread_t thread, int gpuIdx)
{
    std::vector<int> prefer;
    char busId[32] = { 0 };
    if (gpuIdx >= 0 && cudaDeviceGetPCIBusId(busId, sizeof(busId), gpuIdx) ==
                         cudaSuccess) {
        std::string id(busId);
        for (auto& c : id) {
            c = (char)tolower(c);
        }
        std::ifstream f("/sys/bus/pci/devices/" + id + "/local_cpulist");
        std::string line;
        if (f && std::getline(f, line)) {
            prefer = parseCpuList(line);
        }
    } else {
        cudaGetLastError();
    }
    return pinToOneOf(thread, prefer);
}

int gpuForRank(int rank)
{
    int n = getUsableGpus();
    if (n <= 0) {
        return -1;
    }
    const std::string& list = getSystemConfig().gpus;
    if (!list.empty()) {
        std::vector<int> ids = parseCpuList(list);
        if (!ids.empty()) {
            return ids[(size_t)rank % ids.size()];
        }
    }
    return rank % n;
}

void bindThreadToGpu(int gpuIdx)
{
    if (gpuIdx < 0) {
        return;
    }
    if (cudaSetDevice(gpuIdx) != cudaSuccess) {
        cudaGetLastError();
        SPDLOG_WARN("Could not bind thread to GPU {}", gpuIdx);
    }
}

} // namespace faabric::util
