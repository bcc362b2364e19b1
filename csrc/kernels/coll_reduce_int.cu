// Integer MAX / MIN instantiations (signedness matters) + integer dispatch.
// Signed/unsigned SUM and PROD are bit-identical in two's complement and the
// logical/bitwise ops ignore signedness, so those share unsigned instances
// (coll_reduce_int_arith.cu, coll_reduce_int_logic.cu).
#include "coll_reduce.cuh"

namespace fb {

template<typename T, int OP>
static const ReduceLaunchers* L()
{
    return launchersFor<VecReduce<T, OP, false>>();
}

template<int OP>
static const ReduceLaunchers* bySignedType(int dtype)
{
    switch (dtype) {
        case FB_I8:
            return L<int8_t, OP>();
        case FB_U8:
            return L<uint8_t, OP>();
        case FB_I16:
            return L<int16_t, OP>();
        case FB_U16:
            return L<uint16_t, OP>();
        case FB_I32:
            return L<int32_t, OP>();
        case FB_U32:
            return L<uint32_t, OP>();
        case FB_I64:
            return L<int64_t, OP>();
        case FB_U64:
            return L<uint64_t, OP>();
        default:
            return nullptr;
    }
}

const ReduceLaunchers* findReduceLaunchersIntArith(int dtype, int op);
const ReduceLaunchers* findReduceLaunchersIntLogic(int dtype, int op);

const ReduceLaunchers* findReduceLaunchersInt(int dtype, int op)
{
    if (dtype < FB_I8 || dtype > FB_U64) {
        return nullptr;
    }
    switch (op) {
        case FB_OP_MAX:
            return bySignedType<FB_OP_MAX>(dtype);
        case FB_OP_MIN:
            return bySignedType<FB_OP_MIN>(dtype);
        case FB_OP_SUM:
        case FB_OP_PROD:
            return findReduceLaunchersIntArith(dtype, op);
        default:
            return findReduceLaunchersIntLogic(dtype, op);
    }
}

} // namespace fb
