// Lazily resolved CUDA *driver* entry points (VMM + multicast) obtained through
// cudaGetDriverEntryPoint, so libfaabric_b200.so has no link-time dependency on
// libcuda.so.1 and still loads on CPU-only boxes (loopback backend / CI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <string>

namespace faabric::device {

struct DriverApi
{
    bool loaded = false;
    std::string error;

    CUresult (*cuGetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*cuDeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*cuDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) =
      nullptr;
    CUresult (*cuMemGetAllocationGranularity)(
      size_t*,
      const CUmemAllocationProp*,
      CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*cuMemCreate)(CUmemGenericAllocationHandle*,
                            size_t,
                            const CUmemAllocationProp*,
                            unsigned long long) = nullptr;
    CUresult (*cuMemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*cuMemAddressReserve)(CUdeviceptr*,
                                    size_t,
                                    size_t,
                                    CUdeviceptr,
                                    unsigned long long) = nullptr;
    CUresult (*cuMemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*cuMemMap)(CUdeviceptr,
                         size_t,
                         size_t,
                         CUmemGenericAllocationHandle,
                         unsigned long long) = nullptr;
    CUresult (*cuMemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*cuMemSetAccess)(CUdeviceptr,
                               size_t,
                               const CUmemAccessDesc*,
                               size_t) = nullptr;
    CUresult (*cuMemExportToShareableHandle)(void*,
                                             CUmemGenericAllocationHandle,
                                             CUmemAllocationHandleType,
                                             unsigned long long) = nullptr;
    CUresult (*cuMemImportFromShareableHandle)(CUmemGenericAllocationHandle*,
                                               void*,
                                               CUmemAllocationHandleType) =
      nullptr;
    CUresult (*cuMulticastCreate)(CUmemGenericAllocationHandle*,
                                  const CUmulticastObjectProp*) = nullptr;
    CUresult (*cuMulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) =
      nullptr;
    CUresult (*cuMulticastBindMem)(CUmemGenericAllocationHandle,
                                   size_t,
                                   CUmemGenericAllocationHandle,
                                   size_t,
                                   size_t,
                                   unsigned long long) = nullptr;
    CUresult (*cuMulticastGetGranularity)(size_t*,
                                          const CUmulticastObjectProp*,
                                          CUmulticastGranularity_flags) =
      nullptr;

    // stream memory operations (optional): waits that occupy no SM
    CUresult (*cuStreamWaitValue32)(CUstream,
                                    CUdeviceptr,
                                    cuuint32_t,
                                    unsigned int) = nullptr;
    CUresult (*cuStreamWriteValue32)(CUstream,
                                     CUdeviceptr,
                                     cuuint32_t,
                                     unsigned int) = nullptr;

    std::string errStr(CUresult r) const;
};

// Thread-safe, resolves once.  `loaded == false` on CPU-only machines.
const DriverApi& getDriverApi();

// True if a usable CUDA device is present (never throws)
bool cudaAvailable();

int cudaDeviceCountSafe();

} // namespace faabric::device
