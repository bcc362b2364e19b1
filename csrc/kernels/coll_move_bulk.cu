// TMA (bulk async copy engine) variant of the pull collectives for large
// chunks: each CTA streams 32 KiB tiles  peer HBM -> shared memory -> local HBM
// with `cp.async.bulk` and an mbarrier per stage.  One elected thread drives
// the copy engine; no registers or LSU slots are spent on the payload, so a
// handful of CTAs keeps the link busy (which matters when eight lanes share
// the flag budget).  SASS: UBLKCP (load/store), SYNCS.ARRIVE.TRANS64
// (expect_tx), SYNCS.PHASECHK (try_wait).
//
// Same protocol as moveKernel: cross-rank barrier, pull, cross-rank barrier.
#include "coll_move.cuh"

namespace fb {

namespace {
constexpr uint32_t BULK_TILE = 32 * 1024;
constexpr int BULK_STAGES = 4;

__device__ __forceinline__ uint32_t smemPtr(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemPtr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemPtr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t"
                     ".reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t"
                     "}"
                     : "=r"(done)
                     : "r"(smemPtr(bar)), "r"(parity)
                     : "memory");
    }
}

// global (local or peer-mapped) -> shared, completion counted on `bar`
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smemPtr(smemDst)),
                 "l"(gsrc),
                 "r"(bytes),
                 "r"(smemPtr(bar))
                 : "memory");
}

// shared -> global, tracked by the bulk async-group
__device__ __forceinline__ void bulkStore(void* gdst, const void* smemSrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smemPtr(smemSrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

__device__ __forceinline__ void bulkWaitRead0()
{
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__device__ __forceinline__ void bulkWaitAll()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void fenceProxyAsync()
{
    asm volatile("fence.proxy.async;" ::: "memory");
}

struct Piece
{
    const uint8_t* src;
    uint8_t* dst;
};
}

__global__ void __launch_bounds__(128, 1) moveBulkKernel(const MoveArgs a)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[BULK_STAGES];

    BlockBarrier bar;
    bar.load(a.comm);
    bool ok = true;
    if (!a.noSync) {
        ok = bar.sync(a.comm);
    }
    const int rank = a.comm.rank;
    const int n = a.comm.nranks;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BULK_STAGES; s++) {
            mbarInit(&full[s], 1);
        }
        // make the barriers visible to the async proxy, and order the peers'
        // data (acquired through the generic proxy above) before our bulk reads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fenceProxyAsync();
    }
    __syncthreads();

    if (ok && threadIdx.x == 0) {
        // Pieces this rank pulls: (source, destination) of chunkBytes each
        int nPieces = 0;
        int firstPeer = 0;
        if (a.mode == MOVE_ALLGATHER || a.mode == MOVE_ALLTOALL || (a.mode == MOVE_GATHER && rank == a.root)) {
            nPieces = n;
            firstPeer = rank; // start at own rank to spread the load
        } else if (a.mode == MOVE_SCATTER || (a.mode == MOVE_BCAST && rank != a.root)) {
            nPieces = 1;
        }
        const uint64_t srcExtra = (a.mode == MOVE_ALLTOALL) ? (uint64_t)rank * a.srcStride : 0;
        const uint64_t tilesPerPiece = (a.chunkBytes + BULK_TILE - 1) / BULK_TILE;
        const uint64_t totalTiles = (uint64_t)nPieces * tilesPerPiece;

        auto tileSrcDst = [&](uint64_t t, const uint8_t*& src, uint8_t*& dst, uint32_t& bytes) {
            // tile-major interleaving over the pieces keeps all peers busy
            const int q = (int)(t % (uint64_t)max(nPieces, 1));
            const uint64_t k = t / (uint64_t)max(nPieces, 1);
            const uint64_t off = k * BULK_TILE;
            bytes = (uint32_t)min((uint64_t)BULK_TILE, a.chunkBytes - off);
            if (a.mode == MOVE_SCATTER) {
                src = a.comm.heap[a.root] + a.sendOff + (uint64_t)rank * a.srcStride + off;
                dst = a.recvLocal + off;
            } else if (a.mode == MOVE_BCAST) {
                src = a.comm.heap[a.root] + a.sendOff + off;
                dst = a.recvLocal + off;
            } else {
                const int p = (firstPeer + q) % n;
                src = a.comm.heap[p] + a.sendOff + srcExtra + off;
                dst = a.recvLocal + (uint64_t)p * a.dstStride + off;
            }
        };

        // This CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
        const uint64_t first = blockIdx.x;
        const uint64_t step = gridDim.x;
        uint64_t nMine = first < totalTiles ? (totalTiles - first + step - 1) / step : 0;

        // prologue: fill the ring
        for (uint64_t j = 0; j < nMine && j < (uint64_t)BULK_STAGES; j++) {
            const uint8_t* src;
            uint8_t* dst;
            uint32_t bytes;
            tileSrcDst(first + j * step, src, dst, bytes);
            mbarExpectTx(&full[j], bytes);
            bulkLoad(smem + j * BULK_TILE, src, bytes, &full[j]);
        }
        for (uint64_t j = 0; j < nMine; j++) {
            const int s = (int)(j % BULK_STAGES);
            const uint32_t parity = (uint32_t)((j / BULK_STAGES) & 1);
            const uint8_t* src;
            uint8_t* dst;
            uint32_t bytes;
            tileSrcDst(first + j * step, src, dst, bytes);
            mbarWait(&full[s], parity);
            bulkStore(dst, smem + s * BULK_TILE, bytes);
            const uint64_t nxt = j + BULK_STAGES;
            if (nxt < nMine) {
                // the stage can be refilled once the store has read it
                bulkWaitRead0();
                const uint8_t* src2;
                uint8_t* dst2;
                uint32_t bytes2;
                tileSrcDst(first + nxt * step, src2, dst2, bytes2);
                mbarExpectTx(&full[s], bytes2);
                bulkLoad(smem + s * BULK_TILE, src2, bytes2, &full[s]);
            }
        }
        // all writes performed before we tell the peers (and the stream) we are done
        bulkWaitAll();
        fenceProxyAsync();
    }
    __syncthreads();
    if (!a.noSync) {
        bar.sync(a.comm);
    }
    bar.store(a.comm);
}

bool moveBulkSupported(const MoveArgs& a)
{
    switch (a.mode) {
        case MOVE_ALLGATHER:
        case MOVE_ALLTOALL:
        case MOVE_GATHER:
        case MOVE_SCATTER:
        case MOVE_BCAST:
            break;
        default:
            return false;
    }
    return (a.chunkBytes % 16) == 0 && (a.sendOff % 16) == 0 && (a.srcStride % 16) == 0 && (a.dstStride % 16) == 0 &&
           ((uintptr_t)a.recvLocal % 16) == 0;
}

// The opt-in to >48 KiB of dynamic shared memory is per device
static cudaError_t configureBulk()
{
    static bool done[64] = { false };
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        return e;
    }
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(moveBulkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BULK_STAGES * BULK_TILE);
        if (e == cudaSuccess && dev >= 0 && dev < 64) {
            done[dev] = true;
        }
    }
    return e;
}

cudaError_t launchMoveBulk(const MoveArgs& a, int blocks, cudaStream_t s)
{
    cudaError_t e = configureBulk();
    if (e != cudaSuccess) {
        return e;
    }
    moveBulkKernel<<<blocks, 128, BULK_STAGES * BULK_TILE, s>>>(a);
    return cudaGetLastError();
}

cudaError_t preloadMoveBulkKernel()
{
    cudaError_t e = configureBulk();
    if (e != cudaSuccess) {
        return e;
    }
    cudaFuncAttributes attr;
    return cudaFuncGetAttributes(&attr, moveBulkKernel);
}

} // namespace fb
