// One-sided communication for MpiWorld (MPI_Win_*, MPI_Put / MPI_Get).
#include <faabric/mpi/MpiWorld.h>
#include <faabric/util/logging.h>
#include <faabric/util/macros.h>

#include <cuda_runtime.h>

#include <cstring>
#include <stdexcept>

namespace faabric::mpi {

// ---------------------------------------------------------------------------
// One-sided communication
// ---------------------------------------------------------------------------
namespace {
struct RmaSegment
{
    uint64_t base;
    int64_t size;
    int32_t dispUnit;
    int32_t pad;
};

struct RmaWireOp
{
    int32_t kind;
    int32_t pad;
    uint64_t dispBytes;
    uint64_t bytes;
};

void rmaCopy(void* dst, const void* src, size_t bytes)
{
    if (bytes == 0) {
        return;
    }
    if (MpiWorld::isDevicePointer(dst) || MpiWorld::isDevicePointer(src)) {
        if (cudaMemcpy(dst, src, bytes, cudaMemcpyDefault) != cudaSuccess) {
            cudaGetLastError();
            throw std::runtime_error("Device copy for a one-sided operation failed");
        }
    } else {
        memcpy(dst, src, bytes);
    }
}
}

bool MpiWorld::allRanksLocal()
{
    for (int r = 0; r < size; r++) {
        if (!isLocalRank(r)) {
            return false;
        }
    }
    return true;
}

std::shared_ptr<MpiWorld::RmaWindow> MpiWorld::getWindow(int winId)
{
    std::lock_guard<std::mutex> lk(windowsMx);
    auto it = windows.find(winId);
    if (it == windows.end()) {
        SPDLOG_ERROR("MPI window {} does not exist in world {}", winId, id);
        throw std::runtime_error("Unknown MPI window");
    }
    return it->second;
}

int MpiWorld::winCreate(int rank, void* base, int64_t sizeBytes, int dispUnit)
{
    checkRanksRange(0, rank);
    if (sizeBytes < 0 || dispUnit <= 0) {
        throw std::invalid_argument("Bad size / displacement unit for an MPI window");
    }
    int winId = 0;
    std::shared_ptr<RmaWindow> w;
    {
        std::lock_guard<std::mutex> lk(windowsMx);
        if ((int)windowsCreated.size() < size) {
            windowsCreated.resize(size, 0);
        }
        winId = ++windowsCreated[rank];
        auto& slot = windows[winId];
        if (slot == nullptr) {
            slot = std::make_shared<RmaWindow>();
            slot->pending.resize(size);
        }
        w = slot;
    }
    // Everybody learns everybody's segment
    RmaSegment mine{ (uint64_t)(uintptr_t)base, sizeBytes, dispUnit, 0 };
    std::vector<RmaSegment> all(size);
    faabric_datatype_t* byteType = getFaabricDatatypeFromId(FAABRIC_BYTE);
    allGather(rank, BYTES(&mine), byteType, sizeof(RmaSegment), BYTES(all.data()), byteType, sizeof(RmaSegment));
    {
        std::lock_guard<std::mutex> lk(w->mx);
        if (!w->filled) {
            w->bases.resize(size);
            w->sizes.resize(size);
            w->dispUnits.resize(size);
            for (int r = 0; r < size; r++) {
                w->bases[r] = all[r].base;
                w->sizes[r] = all[r].size;
                w->dispUnits[r] = all[r].dispUnit;
            }
            w->filled = true;
        }
    }
    // Nobody may target a segment before its owner has published it
    barrier(rank);
    return winId;
}

void MpiWorld::winFree(int rank, int winId)
{
    auto w = getWindow(winId);
    // Outstanding operations complete first
    winFence(rank, winId);
    int localRanks = 0;
    for (int r = 0; r < size; r++) {
        localRanks += isLocalRank(r) ? 1 : 0;
    }
    bool last = false;
    {
        std::lock_guard<std::mutex> lk(w->mx);
        last = ++w->freed == localRanks;
    }
    if (last) {
        std::lock_guard<std::mutex> lk(windowsMx);
        windows.erase(winId);
    }
}

bool MpiWorld::winQuery(int winId, int rank, void** base, int64_t* sizeBytes, int* dispUnit)
{
    std::shared_ptr<RmaWindow> w;
    {
        std::lock_guard<std::mutex> lk(windowsMx);
        auto it = windows.find(winId);
        if (it == windows.end()) {
            return false;
        }
        w = it->second;
    }
    if (rank < 0 || rank >= size || !w->filled) {
        return false;
    }
    *base = (void*)(uintptr_t)w->bases[rank];
    *sizeBytes = w->sizes[rank];
    *dispUnit = w->dispUnits[rank];
    return true;
}

uint8_t* MpiWorld::winTargetPtr(RmaWindow& w, int targetRank, int64_t targetDisp, size_t bytes)
{
    if (targetRank < 0 || targetRank >= size) {
        throw std::runtime_error("One-sided operation on a rank outside the world");
    }
    const int64_t off = targetDisp * (int64_t)w.dispUnits[targetRank];
    if (targetDisp < 0 || off + (int64_t)bytes > w.sizes[targetRank]) {
        SPDLOG_ERROR("One-sided access [{}, {}) outside the {}-byte window of rank {}", off, off + (int64_t)bytes, w.sizes[targetRank], targetRank);
        throw std::runtime_error("One-sided operation outside the target window");
    }
    return (uint8_t*)(uintptr_t)w.bases[targetRank] + off;
}

void MpiWorld::winPut(int rank, int winId, const uint8_t* origin, size_t bytes, int targetRank, int64_t targetDisp)
{
    auto w = getWindow(winId);
    uint8_t* dst = winTargetPtr(*w, targetRank, targetDisp, bytes);
    if (isLocalRank(targetRank)) {
        // Same address space (or peer-mapped HBM): write it now, the closing
        // fence publishes it
        rmaCopy(dst, origin, bytes);
        return;
    }
    uint64_t dispBytes = (uint64_t)(dst - (uint8_t*)(uintptr_t)w->bases[targetRank]);
    w->pending[rank].push_back(RmaOp{ 0, targetRank, dispBytes, bytes, const_cast<uint8_t*>(origin) });
}

void MpiWorld::winGet(int rank, int winId, uint8_t* origin, size_t bytes, int targetRank, int64_t targetDisp)
{
    auto w = getWindow(winId);
    uint8_t* src = winTargetPtr(*w, targetRank, targetDisp, bytes);
    if (isLocalRank(targetRank)) {
        rmaCopy(origin, src, bytes);
        return;
    }
    uint64_t dispBytes = (uint64_t)(src - (uint8_t*)(uintptr_t)w->bases[targetRank]);
    w->pending[rank].push_back(RmaOp{ 1, targetRank, dispBytes, bytes, origin });
}

void MpiWorld::rmaSendOps(RmaWindow& w, int rank, int peer)
{
    faabric_datatype_t* byteType = getFaabricDatatypeFromId(FAABRIC_BYTE);
    std::vector<const RmaOp*> gets;
    for (const RmaOp& op : w.pending[rank]) {
        if (op.target != peer) {
            continue;
        }
        if (op.bytes > (uint64_t)INT32_MAX) {
            throw std::runtime_error("One-sided operation larger than 2 GiB to another process");
        }
        RmaWireOp wire{ op.kind, 0, op.dispBytes, op.bytes };
        send(rank, peer, BYTES(&wire), byteType, sizeof(wire), MpiMessageType::RMA_OP);
        if (op.kind == 0) {
            send(rank, peer, op.origin, byteType, (int)op.bytes, MpiMessageType::RMA_DATA);
        } else {
            gets.push_back(&op);
        }
    }
    // The target answers each get as it meets it: same order
    for (const RmaOp* op : gets) {
        recv(peer, rank, op->origin, byteType, (int)op->bytes, nullptr, MpiMessageType::RMA_DATA);
    }
}

void MpiWorld::rmaRecvOps(RmaWindow& w, int rank, int peer, int nOps)
{
    faabric_datatype_t* byteType = getFaabricDatatypeFromId(FAABRIC_BYTE);
    uint8_t* base = (uint8_t*)(uintptr_t)w.bases[rank];
    for (int i = 0; i < nOps; i++) {
        RmaWireOp wire{};
        recv(peer, rank, BYTES(&wire), byteType, sizeof(wire), nullptr, MpiMessageType::RMA_OP);
        if ((int64_t)(wire.dispBytes + wire.bytes) > w.sizes[rank]) {
            throw std::runtime_error("Remote one-sided operation outside this rank's window");
        }
        if (wire.kind == 0) {
            recv(peer, rank, base + wire.dispBytes, byteType, (int)wire.bytes, nullptr, MpiMessageType::RMA_DATA);
        } else {
            send(rank, peer, base + wire.dispBytes, byteType, (int)wire.bytes, MpiMessageType::RMA_DATA);
        }
    }
}

void MpiWorld::winFence(int rank, int winId)
{
    auto w = getWindow(winId);
    if (!allRanksLocal()) {
        // How many operations does everybody have for everybody else?
        std::vector<int> outgoing(size, 0), incoming(size, 0);
        for (const RmaOp& op : w->pending[rank]) {
            outgoing[op.target]++;
        }
        faabric_datatype_t* intType = getFaabricDatatypeFromId(FAABRIC_INT);
        allToAll(rank, BYTES(outgoing.data()), intType, 1, BYTES(incoming.data()), intType, 1);
        // Pairwise exchanges in increasing peer order; inside a pair the lower
        // rank ships first.  Every wait is on a strictly "earlier" pair, so the
        // schedule cannot cycle.
        for (int peer = 0; peer < size; peer++) {
            if (peer == rank || isLocalRank(peer)) {
                continue;
            }
            if (rank < peer) {
                rmaSendOps(*w, rank, peer);
                rmaRecvOps(*w, rank, peer, incoming[peer]);
            } else {
                rmaRecvOps(*w, rank, peer, incoming[peer]);
                rmaSendOps(*w, rank, peer);
            }
        }
        w->pending[rank].clear();
    }
    barrier(rank);
}

}
