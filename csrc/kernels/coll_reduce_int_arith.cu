// Integer SUM / PROD instantiations (per width, unsigned lanes)
#include "coll_reduce.cuh"

namespace fb {

template<typename T, int OP>
static const ReduceLaunchers* L()
{
    return launchersFor<VecReduce<T, OP, false>>();
}

template<int OP>
static const ReduceLaunchers* byWidthUnsigned(int dtype)
{
    switch (fbDtypeSize(dtype)) {
        case 1:
            return L<uint8_t, OP>();
        case 2:
            return L<uint16_t, OP>();
        case 4:
            return L<uint32_t, OP>();
        case 8:
            return L<uint64_t, OP>();
        default:
            return nullptr;
    }
}

const ReduceLaunchers* findReduceLaunchersIntArith(int dtype, int op)
{
    if (op == FB_OP_SUM) {
        return byWidthUnsigned<FB_OP_SUM>(dtype);
    }
    if (op == FB_OP_PROD) {
        return byWidthUnsigned<FB_OP_PROD>(dtype);
    }
    return nullptr;
}

} // namespace fb
