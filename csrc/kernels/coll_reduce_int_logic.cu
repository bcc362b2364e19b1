// Integer logical / bitwise instantiations (per width, unsigned lanes)
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

const ReduceLaunchers* findReduceLaunchersIntLogic(int dtype, int op)
{
    switch (op) {
        case FB_OP_LAND:
            return byWidthUnsigned<FB_OP_LAND>(dtype);
        case FB_OP_LOR:
            return byWidthUnsigned<FB_OP_LOR>(dtype);
        case FB_OP_LXOR:
            return byWidthUnsigned<FB_OP_LXOR>(dtype);
        case FB_OP_BAND:
            return byWidthUnsigned<FB_OP_BAND>(dtype);
        case FB_OP_BOR:
            return byWidthUnsigned<FB_OP_BOR>(dtype);
        case FB_OP_BXOR:
            return byWidthUnsigned<FB_OP_BXOR>(dtype);
        default:
            return nullptr;
    }
}

} // namespace fb
