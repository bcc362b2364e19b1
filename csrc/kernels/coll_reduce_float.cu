// Floating-point instantiations (fp32, fp64, fp16, bf16) of the fused
// gather-reduce kernel.  16-bit types combine in fp32.
#include "coll_reduce.cuh"

namespace fb {

template<typename T, int OP>
static const ReduceLaunchers* L()
{
    return launchersFor<VecReduce<T, OP, false>>();
}

template<int OP>
static const ReduceLaunchers* byType(int dtype)
{
    switch (dtype) {
        case FB_F32:
            return L<float, OP>();
        case FB_F64:
            return L<double, OP>();
        case FB_F16:
            return L<__half, OP>();
        case FB_BF16:
            return L<__nv_bfloat16, OP>();
        default:
            return nullptr;
    }
}

const ReduceLaunchers* findReduceLaunchersFloat(int dtype, int op)
{
    switch (op) {
        case FB_OP_MAX:
            return byType<FB_OP_MAX>(dtype);
        case FB_OP_MIN:
            return byType<FB_OP_MIN>(dtype);
        case FB_OP_SUM:
            return byType<FB_OP_SUM>(dtype);
        case FB_OP_PROD:
            return byType<FB_OP_PROD>(dtype);
        default:
            return nullptr;
    }
}

} // namespace fb
