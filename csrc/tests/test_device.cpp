// GPU tests of the host runtime's device paths: MPI on device buffers through
// the fused P2P kernels, device snapshots and device-resident state.
// All are skipped when no CUDA device is visible.
#include "fixtures.h"

#include <faabric/device/cuda_driver.h>
#include <faabric/mpi/MpiWorld.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/mpi/mpi.h>
#include <faabric/snapshot/DeviceSnapshot.h>
#include <faabric/state/State.h>
#include <faabric/util/memory.h>
#include <faabric/util/snapshot.h>

#include <cuda_runtime.h>

#include <numeric>

using namespace tests;

#define NEED_GPU()                                                             \
    do {                                                                       \
        if (!faabric::device::cudaAvailable()) {                               \
            SKIP_TEST("no CUDA device");                                       \
        }                                                                      \
    } while (0)

#define CHECK_RANK(cond)                                                       \
    do {                                                                       \
        if (!(cond)) {                                                         \
            printf("         rank %d: check failed at line %d: %s\n", rank, __LINE__, #cond); \
            return 1;                                                          \
        }                                                                      \
    } while (0)

namespace {
template<typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n;

    explicit DevBuf(size_t nIn)
      : n(nIn)
    {
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
            throw std::runtime_error("cudaMalloc failed");
        }
    }

    ~DevBuf() { cudaFree(p); }

    void upload(const std::vector<T>& v) { cudaMemcpy(p, v.data(), n * sizeof(T), cudaMemcpyHostToDevice); }

    std::vector<T> download()
    {
        std::vector<T> v(n);
        cudaMemcpy(v.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost);
        return v;
    }
};

static std::atomic<uint64_t> lastDeviceCollectives{ 0 };

int deviceCollectivesBody(int rank, int size)
{
    // The executor thread is bound to its GPU by the runtime; collectives pick
    // the rank's device themselves, plain allocations follow the comm
    auto& world = faabric::mpi::getMpiWorldRegistry().getWorld(faabric::executor::ExecutorContext::get()->getMsg().mpiworldid());
    auto comm = world.getDeviceComm(rank);
    CHECK_RANK(comm != nullptr);
    cudaSetDevice(comm->device());

    // All-reduce: int32 sum (the headline op), float max, in place
    const size_t n = 300000;
    DevBuf<int> send(n), recv(n);
    send.upload(std::vector<int>(n, rank + 1));
    MPI_Allreduce(send.p, recv.p, (int)n, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    auto out = recv.download();
    CHECK_RANK(out[0] == size * (size + 1) / 2 && out[n - 1] == out[0]);
    DevBuf<float> f(1000);
    f.upload(std::vector<float>(1000, (float)rank));
    MPI_Allreduce(MPI_IN_PLACE, f.p, 1000, MPI_FLOAT, MPI_MAX, MPI_COMM_WORLD);
    CHECK_RANK(f.download()[999] == (float)(size - 1));

    // Broadcast
    DevBuf<long long> b(5000);
    b.upload(std::vector<long long>(5000, rank == 1 ? 77 : -1));
    MPI_Bcast(b.p, 5000, MPI_LONG_LONG, 1, MPI_COMM_WORLD);
    CHECK_RANK(b.download()[4999] == 77);

    // Allgather / alltoall
    DevBuf<int> mine(64), all(64 * (size_t)size);
    mine.upload(std::vector<int>(64, rank * 3));
    MPI_Allgather(mine.p, 64, MPI_INT, all.p, 64, MPI_INT, MPI_COMM_WORLD);
    auto gathered = all.download();
    for (int r = 0; r < size; r++) {
        CHECK_RANK(gathered[(size_t)r * 64 + 5] == r * 3);
    }
    std::vector<int> a2a((size_t)size * 16);
    for (int r = 0; r < size; r++) {
        std::fill(a2a.begin() + r * 16, a2a.begin() + (r + 1) * 16, rank * 100 + r);
    }
    DevBuf<int> a2aSend(a2a.size()), a2aRecv(a2a.size());
    a2aSend.upload(a2a);
    MPI_Alltoall(a2aSend.p, 16, MPI_INT, a2aRecv.p, 16, MPI_INT, MPI_COMM_WORLD);
    auto exchanged = a2aRecv.download();
    for (int r = 0; r < size; r++) {
        CHECK_RANK(exchanged[(size_t)r * 16 + 3] == r * 100 + rank);
    }

    // Reduce / scan / reduce-scatter / gather / scatter on device memory
    DevBuf<double> d(100), dOut(100);
    d.upload(std::vector<double>(100, rank + 1.0));
    MPI_Reduce(d.p, dOut.p, 100, MPI_DOUBLE, MPI_SUM, 0, MPI_COMM_WORLD);
    if (rank == 0) {
        CHECK_RANK(dOut.download()[50] == size * (size + 1) / 2.0);
    }
    MPI_Scan(d.p, dOut.p, 100, MPI_DOUBLE, MPI_SUM, MPI_COMM_WORLD);
    CHECK_RANK(dOut.download()[0] == (rank + 1) * (rank + 2) / 2.0);
    DevBuf<int> rsIn(8 * (size_t)size), rsOut(8);
    rsIn.upload(std::vector<int>(8 * (size_t)size, rank));
    std::vector<int> counts(size, 8);
    MPI_Reduce_scatter(rsIn.p, rsOut.p, counts.data(), MPI_INT, MPI_SUM, MPI_COMM_WORLD);
    CHECK_RANK(rsOut.download()[7] == size * (size - 1) / 2);

    // MPI_IN_PLACE at the root, device buffers: only the root can see that it
    // is in place, yet every rank must end up on the same (device) path
    DevBuf<int> ip(500);
    ip.upload(std::vector<int>(500, rank + 1));
    if (rank == 0) {
        MPI_Reduce(MPI_IN_PLACE, ip.p, 500, MPI_INT, MPI_SUM, 0, MPI_COMM_WORLD);
        auto red = ip.download();
        CHECK_RANK(red[0] == size * (size + 1) / 2 && red[499] == red[0]);
    } else {
        MPI_Reduce(ip.p, nullptr, 500, MPI_INT, MPI_SUM, 0, MPI_COMM_WORLD);
    }
    const int gRoot = size - 1;
    DevBuf<int> gAll(64 * (size_t)size);
    if (rank == gRoot) {
        std::vector<int> pre(64 * (size_t)size, -7);
        std::fill(pre.begin() + (size_t)gRoot * 64, pre.begin() + (size_t)(gRoot + 1) * 64, gRoot * 3);
        gAll.upload(pre);
        MPI_Gather(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, gAll.p, 64, MPI_INT, gRoot, MPI_COMM_WORLD);
        auto gg = gAll.download();
        for (int r = 0; r < size; r++) {
            CHECK_RANK(gg[(size_t)r * 64] == r * 3 && gg[(size_t)r * 64 + 63] == r * 3);
        }
    } else {
        MPI_Gather(mine.p, 64, MPI_INT, nullptr, 0, MPI_INT, gRoot, MPI_COMM_WORLD);
    }

    // Point to point with device buffers (both ends on the device, and mixed)
    int right = (rank + 1) % size, left = (rank + size - 1) % size;
    DevBuf<int> tok(2048), got(2048);
    tok.upload(std::vector<int>(2048, rank));
    MPI_Sendrecv(tok.p, 2048, MPI_INT, right, 0, got.p, 2048, MPI_INT, left, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    CHECK_RANK(got.download()[2047] == left);
    std::vector<int> hostGot(2048, -1);
    MPI_Sendrecv(tok.p, 2048, MPI_INT, right, 0, hostGot.data(), 2048, MPI_INT, left, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
    CHECK_RANK(hostGot[0] == left);

    MPI_Barrier(MPI_COMM_WORLD);
    if (rank == 0) {
        lastDeviceCollectives = world.getDeviceCollectiveCount();
    }
    return 0;
}

void runDeviceMpi(int worldSize)
{
    ClusterFixture f(worldSize);
    registerTestFunction("mpi", "device-collectives", [&](auto*, int, int, auto) {
        MPI_Init(nullptr, nullptr);
        int rank = -1, size = -1;
        MPI_Comm_rank(MPI_COMM_WORLD, &rank);
        MPI_Comm_size(MPI_COMM_WORLD, &size);
        int rc = deviceCollectivesBody(rank, size);
        MPI_Finalize();
        return rc;
    });
    auto req = faabric::util::batchExecFactory("mpi", "device-collectives", 1);
    req->mutable_messages(0)->set_ismpi(true);
    req->mutable_messages(0)->set_mpiworldsize(worldSize);
    lastDeviceCollectives = 0;
    f.plannerCli.callFunctions(req);
    auto status = f.awaitBatch(req, 120000);
    REQUIRE_EQ(status->messageresults_size(), worldSize);
    for (auto& m : status->messageresults()) {
        if (m.returnvalue() != 0) {
            fbtest::fail(__FILE__, __LINE__, "rank " + std::to_string(m.mpirank()) + " failed: " + m.outputdata());
        }
    }
    // The collectives really ran as fused device kernels, not via host staging
    printf("         device collectives run by the world: %lu\n", (unsigned long)lastDeviceCollectives.load());
    REQUIRE(lastDeviceCollectives.load() >= (uint64_t)worldSize * 8);
    faabric::mpi::getMpiWorldRegistry().clear();
}
}

TEST_CASE("gpu: MPI collectives on device buffers, 2 ranks", "[gpu][mpi]")
{
    NEED_GPU();
    runDeviceMpi(2);
}

TEST_CASE("gpu: MPI collectives on device buffers, one rank per GPU (or 4 sharing)", "[gpu][mpi]")
{
    NEED_GPU();
    int n = faabric::device::cudaDeviceCountSafe();
    runDeviceMpi(n >= 4 ? n : 4);
}

TEST_CASE("gpu: device snapshot diff+merge+push matches the host implementation", "[gpu][snapshot]")
{
    NEED_GPU();
    cudaSetDevice(0);
    using namespace faabric::util;
    getSystemConfig().diffingMode = "bytewise";
    const size_t size = 64 * HOST_PAGE_SIZE;
    std::vector<uint8_t> base(size);
    for (size_t i = 0; i < size; i++) {
        base[i] = (uint8_t)(i * 13 + 1);
    }
    int baseInt = 1000;
    double baseDouble = 1.5;
    memcpy(base.data() + 4096 + 16, &baseInt, sizeof(int));
    memcpy(base.data() + 8192 + 64, &baseDouble, sizeof(double));

    // Worker-side memory after some work
    std::vector<uint8_t> updated = base;
    int newInt = 1042;
    double newDouble = 6.0;
    memcpy(updated.data() + 4096 + 16, &newInt, sizeof(int));
    memcpy(updated.data() + 8192 + 64, &newDouble, sizeof(double));
    for (size_t i = 20000; i < 20100; i++) {
        updated[i] ^= 0x5a;
    }
    updated[size - 1] = 0;
    updated[12288 + 5] = 99; // inside an ignored region

    auto addRegions = [](auto& snap) {
        snap.addMergeRegion(4096 + 16, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum);
        snap.addMergeRegion(8192 + 64, sizeof(double), SnapshotDataType::Double, SnapshotMergeOperation::Product);
        snap.addMergeRegion(12288, 256, SnapshotDataType::Raw, SnapshotMergeOperation::Ignore);
    };

    // Host reference: diff then apply on a "main" copy that already has
    // another worker's contribution
    std::vector<uint8_t> mainStart = base;
    int otherInt = 1005;
    memcpy(mainStart.data() + 4096 + 16, &otherInt, sizeof(int));
    SnapshotData hostSnap(std::span<const uint8_t>(base.data(), size));
    addRegions(hostSnap);
    hostSnap.fillGapsWithBytewiseRegions();
    // NB the host diff rewrites typed values in the memory it is given with
    // their deltas (reference semantics), so it works on a scratch copy
    std::vector<uint8_t> hostScratch = updated;
    auto diffs = hostSnap.diffWithDirtyRegions(std::span<uint8_t>(hostScratch.data(), size), std::vector<char>(64, 1));
    SnapshotData hostMain(std::span<const uint8_t>(mainStart.data(), size));
    hostMain.applyDiffs(diffs);
    auto expected = hostMain.getDataCopy();

    // Device: one fused kernel does diff + typed merge + push into main
    faabric::snapshot::DeviceSnapshot devSnap(size, 0);
    devSnap.copyInData(base);
    addRegions(devSnap);
    faabric::snapshot::DeviceSnapshot devMain(size, 0);
    devMain.copyInData(mainStart);
    auto mem = allocateDeviceMemory(size, 0);
    cudaMemcpy(mem.ptr, updated.data(), size, cudaMemcpyHostToDevice);
    devSnap.diffAndPush(mem.ptr, size, devMain.getDevicePtr(), nullptr, false, nullptr);
    auto stats = devSnap.getLastStats(nullptr);
    auto got = devMain.getDataCopy(0, size);
    size_t mismatches = 0;
    for (size_t i = 0; i < size; i++) {
        if (got[i] != expected[i]) {
            if (mismatches < 8) {
                printf("         mismatch at %zu: device %d host %d (base %d updated %d)\n", i, got[i], expected[i], base[i], updated[i]);
            }
            mismatches++;
        }
    }
    REQUIRE_EQ(mismatches, 0u);
    REQUIRE(stats.diffBytes > 0);
    int mergedInt;
    memcpy(&mergedInt, got.data() + 4096 + 16, sizeof(int));
    REQUIRE_EQ(mergedInt, 1005 + 42);
    REQUIRE_EQ(got[12288 + 5], base[12288 + 5]);

    // The dirty-page scan agrees with a host comparison
    auto dirty = devSnap.getDirtyPages(mem.ptr, size);
    REQUIRE_EQ(dirty.size(), 64u);
    for (size_t p = 0; p < 64; p++) {
        bool differs = memcmp(base.data() + p * HOST_PAGE_SIZE, updated.data() + p * HOST_PAGE_SIZE, HOST_PAGE_SIZE) != 0;
        REQUIRE_EQ((bool)dirty[p], differs);
    }

    // Host-produced diffs can be applied to a device image
    faabric::snapshot::DeviceSnapshot devMain2(size, 0);
    devMain2.copyInData(mainStart);
    devMain2.applyDiffs(diffs, nullptr);
    REQUIRE(devMain2.getDataCopy(0, size) == expected);
    getSystemConfig().reset();
}

TEST_CASE("gpu: state values have a coherent device copy", "[gpu][state]")
{
    NEED_GPU();
    cudaSetDevice(0);
    auto& state = faabric::state::getGlobalState();
    state.forceClearAll(true);
    size_t size = 3 * STATE_STREAMING_CHUNK_SIZE + 100;
    auto kv = state.getKV("demo", "devstate", size);
    std::vector<uint8_t> init(size, 4);
    kv->set(init.data());
    uint8_t* dev = kv->getDevicePtr(0);
    REQUIRE(dev != nullptr);
    REQUIRE(kv->hasDeviceCopy());
    cudaDeviceSynchronize();
    std::vector<uint8_t> check(16);
    cudaMemcpy(check.data(), dev + STATE_STREAMING_CHUNK_SIZE, 16, cudaMemcpyDeviceToHost);
    REQUIRE_EQ(check[0], 4);
    // Device-side update flows back to the host value
    cudaMemset(dev + 2 * STATE_STREAMING_CHUNK_SIZE + 10, 9, 50);
    kv->flagDeviceChunkDirty(2 * STATE_STREAMING_CHUNK_SIZE + 10, 50);
    kv->syncFromDevice();
    REQUIRE_EQ(*kv->getChunk(2 * STATE_STREAMING_CHUNK_SIZE + 10, 1), 9);
    REQUIRE_EQ(*kv->getChunk(2 * STATE_STREAMING_CHUNK_SIZE + 9, 1), 4);
    // Host-side update reaches the device on the next getDevicePtr
    uint8_t v = 7;
    kv->setChunk(5, &v, 1);
    dev = kv->getDevicePtr(0);
    cudaDeviceSynchronize();
    cudaMemcpy(check.data(), dev, 16, cudaMemcpyDeviceToHost);
    REQUIRE_EQ(check[5], 7);
    REQUIRE_EQ(check[6], 4);
    state.forceClearAll(true);
}

namespace {
int symmetricNonBlockingBody(int rank, int size)
{
    auto& world = faabric::mpi::getMpiWorldRegistry().getWorld(faabric::executor::ExecutorContext::get()->getMsg().mpiworldid());
    auto comm = world.getDeviceComm(rank);
    CHECK_RANK(comm != nullptr);
    cudaSetDevice(comm->device());
    // MPI_Alloc_mem(MPI_INFO_FAABRIC_DEVICE): symmetric heap, no staging
    const int nTensors = 40;
    const size_t per = 5000;
    int *send = nullptr, *recv = nullptr;
    CHECK_RANK(MPI_Alloc_mem(nTensors * per * sizeof(int), MPI_INFO_FAABRIC_DEVICE, &send) == MPI_SUCCESS);
    CHECK_RANK(MPI_Alloc_mem(nTensors * per * sizeof(int), MPI_INFO_FAABRIC_DEVICE, &recv) == MPI_SUCCESS);
    CHECK_RANK(comm->inHeap(send, nTensors * per * sizeof(int)));
    std::vector<int> init(nTensors * per);
    for (size_t i = 0; i < init.size(); i++) {
        init[i] = (int)(i / per) + rank;
    }
    cudaMemcpy(send, init.data(), init.size() * sizeof(int), cudaMemcpyHostToDevice);
    // A burst of non-blocking all-reduces pipelines over the channels
    std::vector<MPI_Request> reqs(nTensors);
    for (int round = 0; round < 3; round++) {
        for (int t = 0; t < nTensors; t++) {
            MPI_Iallreduce(send + t * per, recv + t * per, (int)per, MPI_INT, MPI_SUM, MPI_COMM_WORLD, &reqs[t]);
        }
        MPI_Waitall(nTensors, reqs.data(), MPI_STATUSES_IGNORE);
        std::vector<int> out(nTensors * per);
        cudaMemcpy(out.data(), recv, out.size() * sizeof(int), cudaMemcpyDeviceToHost);
        for (int t = 0; t < nTensors; t++) {
            int expected = t * size + size * (size - 1) / 2;
            CHECK_RANK(out[t * per] == expected && out[(t + 1) * per - 1] == expected);
        }
    }
    // Blocking calls on symmetric memory still work, as do host buffers
    MPI_Allreduce(send, recv, (int)per, MPI_INT, MPI_MAX, MPI_COMM_WORLD);
    int hostVal = rank, hostMax = -1;
    MPI_Request hostReq;
    MPI_Iallreduce(&hostVal, &hostMax, 1, MPI_INT, MPI_MAX, MPI_COMM_WORLD, &hostReq);
    MPI_Wait(&hostReq, MPI_STATUS_IGNORE);
    CHECK_RANK(hostMax == size - 1);
    MPI_Free_mem(send);
    MPI_Free_mem(recv);
    MPI_Barrier(MPI_COMM_WORLD);
    return 0;
}
}

TEST_CASE("gpu: MPI_Iallreduce bursts on MPI_Alloc_mem device memory", "[gpu][mpi]")
{
    NEED_GPU();
    int worldSize = std::max(2, faabric::device::cudaDeviceCountSafe());
    ClusterFixture f(worldSize);
    registerTestFunction("mpi", "device-iallreduce", [&](auto*, int, int, auto) {
        MPI_Init(nullptr, nullptr);
        int rank = -1, size = -1;
        MPI_Comm_rank(MPI_COMM_WORLD, &rank);
        MPI_Comm_size(MPI_COMM_WORLD, &size);
        int rc = symmetricNonBlockingBody(rank, size);
        MPI_Finalize();
        return rc;
    });
    auto req = faabric::util::batchExecFactory("mpi", "device-iallreduce", 1);
    req->mutable_messages(0)->set_ismpi(true);
    req->mutable_messages(0)->set_mpiworldsize(worldSize);
    f.plannerCli.callFunctions(req);
    auto status = f.awaitBatch(req, 120000);
    REQUIRE_EQ(status->messageresults_size(), worldSize);
    for (auto& m : status->messageresults()) {
        if (m.returnvalue() != 0) {
            fbtest::fail(__FILE__, __LINE__, "rank " + std::to_string(m.mpirank()) + " failed: " + m.outputdata());
        }
    }
    faabric::mpi::getMpiWorldRegistry().clear();
}
