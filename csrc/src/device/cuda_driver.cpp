#include "faabric/device/cuda_driver.h"

#include <mutex>

namespace faabric::device {

std::string DriverApi::errStr(CUresult r) const
{
    const char* s = nullptr;
    if (cuGetErrorString != nullptr && cuGetErrorString(r, &s) == CUDA_SUCCESS &&
        s != nullptr) {
        return std::string(s) + " (" + std::to_string((int)r) + ")";
    }
    return "CUresult " + std::to_string((int)r);
}

template<typename F>
static bool resolve(DriverApi& api, const char* name, F& out)
{
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || fn == nullptr ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        api.error += std::string(name) + " ";
        return false;
    }
    out = reinterpret_cast<F>(fn);
    return true;
}

static DriverApi loadApi()
{
    DriverApi api;
    if (!cudaAvailable()) {
        api.error = "no CUDA device";
        return api;
    }
    bool ok = true;
#define FB_RESOLVE(n) ok = resolve(api, #n, api.n) && ok
    FB_RESOLVE(cuGetErrorString);
    FB_RESOLVE(cuDeviceGet);
    FB_RESOLVE(cuDeviceGetAttribute);
    FB_RESOLVE(cuMemGetAllocationGranularity);
    FB_RESOLVE(cuMemCreate);
    FB_RESOLVE(cuMemRelease);
    FB_RESOLVE(cuMemAddressReserve);
    FB_RESOLVE(cuMemAddressFree);
    FB_RESOLVE(cuMemMap);
    FB_RESOLVE(cuMemUnmap);
    FB_RESOLVE(cuMemSetAccess);
    FB_RESOLVE(cuMemExportToShareableHandle);
    FB_RESOLVE(cuMemImportFromShareableHandle);
    bool core = ok;
    // multicast is optional
    FB_RESOLVE(cuMulticastCreate);
    FB_RESOLVE(cuMulticastAddDevice);
    FB_RESOLVE(cuMulticastBindMem);
    FB_RESOLVE(cuMulticastGetGranularity);
    FB_RESOLVE(cuStreamWaitValue32);
    FB_RESOLVE(cuStreamWriteValue32);
#undef FB_RESOLVE
    api.loaded = core;
    return api;
}

const DriverApi& getDriverApi()
{
    static DriverApi api = loadApi();
    return api;
}

int cudaDeviceCountSafe()
{
    static int count = []() {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return n;
    }();
    return count;
}

bool cudaAvailable()
{
    return cudaDeviceCountSafe() > 0;
}

} // namespace faabric::device
