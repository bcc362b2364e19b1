// NVLS (NVSwitch multicast) collectives: the reduction happens *inside the
// switch* via multimem.ld_reduce on a multicast address, and results are
// broadcast with a single multimem.st.  Per-GPU NVLink traffic drops from
// 2(N-1)/N*S (P2P two-shot) to roughly S/N in + S/N out per phase.
// Only a subset of (dtype, op) pairs exists in hardware; everything else uses
// the P2P kernels in coll_reduce*.cu.  SASS: LDGMC.*, STG to multicast VA.
#include "coll_nvls.cuh"

namespace fb {

enum MmVariant
{
    MM_ADD_F32 = 0,
    MM_ADD_F16,
    MM_MIN_F16,
    MM_MAX_F16,
    MM_ADD_BF16,
    MM_MIN_BF16,
    MM_MAX_BF16,
    MM_ADD_U32,
    MM_MIN_U32,
    MM_MAX_U32,
    MM_MIN_S32,
    MM_MAX_S32,
    MM_ADD_U64,
    MM_MIN_U64,
    MM_MAX_U64,
    MM_MIN_S64,
    MM_MAX_S64,
    MM_ADD_F64,
    MM_AND_B32,
    MM_OR_B32,
    MM_XOR_B32,
    MM_COPY // no reduction (bcast / allgather)
};

#define MM_V4(OPSTR)                                                           \
    asm volatile("multimem.ld_reduce.relaxed.sys.global." OPSTR                \
                 " {%0,%1,%2,%3}, [%4];"                                       \
                 : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3])      \
                 : "l"(p)                                                      \
                 : "memory")

#define MM_S32x4(OPSTR)                                                        \
    _Pragma("unroll") for (int k = 0; k < 4; k++)                              \
    {                                                                          \
        asm volatile("multimem.ld_reduce.relaxed.sys.global." OPSTR            \
                     " %0, [%1];"                                              \
                     : "=r"(v.w[k])                                            \
                     : "l"(p + 4 * k)                                          \
                     : "memory");                                              \
    }

#define MM_S64x2(OPSTR)                                                        \
    _Pragma("unroll") for (int k = 0; k < 2; k++)                              \
    {                                                                          \
        uint64_t t;                                                            \
        asm volatile("multimem.ld_reduce.relaxed.sys.global." OPSTR            \
                     " %0, [%1];"                                              \
                     : "=l"(t)                                                 \
                     : "l"(p + 8 * k)                                          \
                     : "memory");                                              \
        v.w[2 * k] = (uint32_t)t;                                              \
        v.w[2 * k + 1] = (uint32_t)(t >> 32);                                  \
    }

template<int V>
__device__ __forceinline__ Vec16 mmLoadReduce(const uint8_t* p)
{
    Vec16 v;
    if constexpr (V == MM_ADD_F32) {
        MM_V4("add.v4.f32");
    } else if constexpr (V == MM_ADD_F16) {
        MM_V4("add.acc::f32.v4.f16x2");
    } else if constexpr (V == MM_MIN_F16) {
        MM_V4("min.v4.f16x2");
    } else if constexpr (V == MM_MAX_F16) {
        MM_V4("max.v4.f16x2");
    } else if constexpr (V == MM_ADD_BF16) {
        MM_V4("add.acc::f32.v4.bf16x2");
    } else if constexpr (V == MM_MIN_BF16) {
        MM_V4("min.v4.bf16x2");
    } else if constexpr (V == MM_MAX_BF16) {
        MM_V4("max.v4.bf16x2");
    } else if constexpr (V == MM_ADD_U32) {
        MM_S32x4("add.u32");
    } else if constexpr (V == MM_MIN_U32) {
        MM_S32x4("min.u32");
    } else if constexpr (V == MM_MAX_U32) {
        MM_S32x4("max.u32");
    } else if constexpr (V == MM_MIN_S32) {
        MM_S32x4("min.s32");
    } else if constexpr (V == MM_MAX_S32) {
        MM_S32x4("max.s32");
    } else if constexpr (V == MM_ADD_U64) {
        MM_S64x2("add.u64");
    } else if constexpr (V == MM_MIN_U64) {
        MM_S64x2("min.u64");
    } else if constexpr (V == MM_MAX_U64) {
        MM_S64x2("max.u64");
    } else if constexpr (V == MM_MIN_S64) {
        MM_S64x2("min.s64");
    } else if constexpr (V == MM_MAX_S64) {
        MM_S64x2("max.s64");
    } else if constexpr (V == MM_ADD_F64) {
        _Pragma("unroll") for (int k = 0; k < 2; k++)
        {
            double t;
            asm volatile(
              "multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];"
              : "=d"(t)
              : "l"(p + 8 * k)
              : "memory");
            uint64_t u = (uint64_t)__double_as_longlong(t);
            v.w[2 * k] = (uint32_t)u;
            v.w[2 * k + 1] = (uint32_t)(u >> 32);
        }
    } else if constexpr (V == MM_AND_B32) {
        MM_S32x4("and.b32");
    } else if constexpr (V == MM_OR_B32) {
        MM_S32x4("or.b32");
    } else if constexpr (V == MM_XOR_B32) {
        MM_S32x4("xor.b32");
    } else {
        v = ldVecStream(p);
    }
    return v;
}

__device__ __forceinline__ void mmStore(uint8_t* p, const Vec16& v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::
                   "l"(p),
                 "f"(__uint_as_float(v.w[0])),
                 "f"(__uint_as_float(v.w[1])),
                 "f"(__uint_as_float(v.w[2])),
                 "f"(__uint_as_float(v.w[3]))
                 : "memory");
}

template<int V>
__global__ void __launch_bounds__(512, 1) nvlsKernel(const NvlsArgs a)
{
    BlockBarrier bar;
    bar.load(a.comm);
    bool ok = a.noSync ? true : bar.sync(a.comm);
    if (ok) {
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        uint64_t i = a.vecBegin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const uint8_t* src;
        if (a.mode == NVLS_BCAST || a.mode == NVLS_ALLGATHER) {
            src = a.comm.heap[a.comm.rank] + a.sendOff; // plain local loads
        } else {
            src = a.comm.mcHeap + a.sendOff; // in-switch reduction
        }
        const bool toAll = (a.mode != NVLS_REDUCE_LOCAL);
        constexpr int U = 4;
        for (; i + (U - 1) * stride < a.vecEnd; i += U * stride) {
            Vec16 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                v[u] = mmLoadReduce<V>(src + (i + u * stride) * 16);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint64_t o = (i + u * stride - a.outBase) * 16;
                if (toAll) {
                    mmStore(a.comm.mcHeap + a.recvOff + o, v[u]);
                } else {
                    stVec(a.recvLocal + o, v[u]);
                }
            }
        }
        for (; i < a.vecEnd; i += stride) {
            Vec16 v = mmLoadReduce<V>(src + i * 16);
            uint64_t o = (i - a.outBase) * 16;
            if (toAll) {
                mmStore(a.comm.mcHeap + a.recvOff + o, v);
            } else {
                stVec(a.recvLocal + o, v);
            }
        }
    }
    if (!a.noSync) {
        bar.sync(a.comm);
    }
    bar.store(a.comm);
}

static int variantFor(int dtype, int op)
{
    switch (dtype) {
        case FB_F32:
            return op == FB_OP_SUM ? MM_ADD_F32 : -1;
        case FB_F64:
            return op == FB_OP_SUM ? MM_ADD_F64 : -1;
        case FB_F16:
            return op == FB_OP_SUM   ? MM_ADD_F16
                   : op == FB_OP_MIN ? MM_MIN_F16
                   : op == FB_OP_MAX ? MM_MAX_F16
                                     : -1;
        case FB_BF16:
            return op == FB_OP_SUM   ? MM_ADD_BF16
                   : op == FB_OP_MIN ? MM_MIN_BF16
                   : op == FB_OP_MAX ? MM_MAX_BF16
                                     : -1;
        case FB_I32:
            return op == FB_OP_SUM   ? MM_ADD_U32
                   : op == FB_OP_MIN ? MM_MIN_S32
                   : op == FB_OP_MAX ? MM_MAX_S32
                                     : -2;
        case FB_U32:
            return op == FB_OP_SUM   ? MM_ADD_U32
                   : op == FB_OP_MIN ? MM_MIN_U32
                   : op == FB_OP_MAX ? MM_MAX_U32
                                     : -2;
        case FB_I64:
            return op == FB_OP_SUM   ? MM_ADD_U64
                   : op == FB_OP_MIN ? MM_MIN_S64
                   : op == FB_OP_MAX ? MM_MAX_S64
                                     : -2;
        case FB_U64:
            return op == FB_OP_SUM   ? MM_ADD_U64
                   : op == FB_OP_MIN ? MM_MIN_U64
                   : op == FB_OP_MAX ? MM_MAX_U64
                                     : -2;
        case FB_I8:
        case FB_U8:
        case FB_I16:
        case FB_U16:
            return -2;
        default:
            return -1;
    }
}

int nvlsVariant(int dtype, int op)
{
    int v = variantFor(dtype, op);
    if (v == -2) {
        // integer type: bitwise ops are width agnostic
        if (op == FB_OP_BAND) {
            return MM_AND_B32;
        }
        if (op == FB_OP_BOR) {
            return MM_OR_B32;
        }
        if (op == FB_OP_BXOR) {
            return MM_XOR_B32;
        }
        return -1;
    }
    return v;
}

// Only the floating-point variants have 16-byte (.v4) multimem forms; integer
// and f64 reductions issue one switch request per 4/8-byte element, which
// makes them request-rate bound for all but very large messages
bool nvlsVectorised(int variant)
{
    switch (variant) {
        case MM_ADD_F32:
        case MM_ADD_F16:
        case MM_MIN_F16:
        case MM_MAX_F16:
        case MM_ADD_BF16:
        case MM_MIN_BF16:
        case MM_MAX_BF16:
        case MM_COPY:
            return true;
        default:
            return false;
    }
}

#define NVLS_CASE(V)                                                           \
    case V:                                                                    \
        nvlsKernel<V><<<blocks, threads, 0, s>>>(a);                           \
        break;

cudaError_t launchNvls(const NvlsArgs& a,
                       int variant,
                       int blocks,
                       int threads,
                       cudaStream_t s)
{
    if (a.mode == NVLS_BCAST || a.mode == NVLS_ALLGATHER) {
        variant = MM_COPY;
    }
    switch (variant) {
        NVLS_CASE(MM_ADD_F32)
        NVLS_CASE(MM_ADD_F16)
        NVLS_CASE(MM_MIN_F16)
        NVLS_CASE(MM_MAX_F16)
        NVLS_CASE(MM_ADD_BF16)
        NVLS_CASE(MM_MIN_BF16)
        NVLS_CASE(MM_MAX_BF16)
        NVLS_CASE(MM_ADD_U32)
        NVLS_CASE(MM_MIN_U32)
        NVLS_CASE(MM_MAX_U32)
        NVLS_CASE(MM_MIN_S32)
        NVLS_CASE(MM_MAX_S32)
        NVLS_CASE(MM_ADD_U64)
        NVLS_CASE(MM_MIN_U64)
        NVLS_CASE(MM_MAX_U64)
        NVLS_CASE(MM_MIN_S64)
        NVLS_CASE(MM_MAX_S64)
        NVLS_CASE(MM_ADD_F64)
        NVLS_CASE(MM_AND_B32)
        NVLS_CASE(MM_OR_B32)
        NVLS_CASE(MM_XOR_B32)
        NVLS_CASE(MM_COPY)
        default:
            return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#define NVLS_PRELOAD(V)                                                        \
    if (e == cudaSuccess) {                                                    \
        e = cudaFuncGetAttributes(&a, nvlsKernel<V>);                          \
    }

cudaError_t preloadNvlsKernels()
{
    cudaFuncAttributes a;
    cudaError_t e = cudaSuccess;
    NVLS_PRELOAD(MM_ADD_F32)
    NVLS_PRELOAD(MM_ADD_F16)
    NVLS_PRELOAD(MM_MIN_F16)
    NVLS_PRELOAD(MM_MAX_F16)
    NVLS_PRELOAD(MM_ADD_BF16)
    NVLS_PRELOAD(MM_MIN_BF16)
    NVLS_PRELOAD(MM_MAX_BF16)
    NVLS_PRELOAD(MM_ADD_U32)
    NVLS_PRELOAD(MM_MIN_U32)
    NVLS_PRELOAD(MM_MAX_U32)
    NVLS_PRELOAD(MM_MIN_S32)
    NVLS_PRELOAD(MM_MAX_S32)
    NVLS_PRELOAD(MM_ADD_U64)
    NVLS_PRELOAD(MM_MIN_U64)
    NVLS_PRELOAD(MM_MAX_U64)
    NVLS_PRELOAD(MM_MIN_S64)
    NVLS_PRELOAD(MM_MAX_S64)
    NVLS_PRELOAD(MM_ADD_F64)
    NVLS_PRELOAD(MM_AND_B32)
    NVLS_PRELOAD(MM_OR_B32)
    NVLS_PRELOAD(MM_XOR_B32)
    NVLS_PRELOAD(MM_COPY)
    return e;
}

} // namespace fb
