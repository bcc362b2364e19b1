// CPU pinning for busy-waiting rank threads + rank -> GPU placement
// (reference: src/util/hwloc.cpp:15-109 pins threads only).
#pragma once

#include <memory>
#include <pthread.h>
#include <sched.h>

namespace faabric::util {

// RAII claim on one CPU of the free-CPU set; released on destruction
class FaabricCpuSet
{
  public:
    explicit FaabricCpuSet(int cpuIdxIn = -1);
    FaabricCpuSet(const FaabricCpuSet&) = delete;
    FaabricCpuSet& operator=(const FaabricCpuSet&) = delete;
    ~FaabricCpuSet();

    cpu_set_t* get() { return &cpuSet; }
    int getCpuIdx() const { return cpuIdx; }

  private:
    cpu_set_t cpuSet;
    int cpuIdx;
};

// Pins the thread to a currently unclaimed CPU (throws if none left)
std::unique_ptr<FaabricCpuSet> pinThreadToFreeCpu(pthread_t thread);

// Pin near a GPU: picks a free CPU from the NUMA node the GPU hangs off when
// that can be determined from sysfs, any free CPU otherwise
std::unique_ptr<FaabricCpuSet> pinThreadNearGpu(pthread_t thread, int gpuIdx);

// Round-robin placement of an MPI rank / executor slot onto the visible GPUs
// (-1 when there is no GPU)
int gpuForRank(int rank);

// Binds the calling thread to a GPU (cudaSetDevice); no-op without GPUs
void bindThreadToGpu(int gpuIdx);

int getNumFreeCpus();

}
