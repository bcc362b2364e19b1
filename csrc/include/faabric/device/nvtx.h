// Optional NVTX ranges around collectives (FAABRIC_NVTX=1): they show up in
// Nsight Systems / ncu timelines.  Header-only NVTX3: no link dependency.
#pragma once

#include <nvtx3/nvToolsExt.h>

#include <cstdlib>

namespace faabric::device {

inline bool nvtxEnabled()
{
    static bool on = []() {
        const char* v = getenv("FAABRIC_NVTX");
        return v != nullptr && v[0] == '1';
    }();
    return on;
}

class NvtxRange
{
  public:
    explicit NvtxRange(const char* name)
      : active(nvtxEnabled())
    {
        if (active) {
            nvtxRangePushA(name);
        }
    }

    ~NvtxRange()
    {
        if (active) {
            nvtxRangePop();
        }
    }

    NvtxRange(const NvtxRange&) = delete;

  private:
    bool active;
};

}
