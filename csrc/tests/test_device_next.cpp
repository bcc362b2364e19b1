// Device-path tests written WITHOUT access to a GPU (the round's GPU budget was
// spent): tag [unverified] keeps them out of `--tag gpu`, which is what the
// pytest GPU suite runs.  First thing to do with a B200: run
//   build/bin/faabric_tests --tag unverified
// fix what they find, then retag them [gpu].  They skip without a device.
#include "fixtures.h"

#include <faabric/device/communicator.h>
#include <faabric/device/cuda_driver.h>
#include <faabric/mpi/MpiWorld.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/mpi/mpi.h>
#include <faabric/snapshot/DeviceSnapshot.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/util/snapshot.h>

#include <cuda_runtime.h>

#include <filesystem>
#include <unistd.h>

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
void runMpiOnGpus(const std::string& name, int worldSize, const std::function<int(int, int)>& body)
{
    ClusterFixture f(worldSize);
    registerTestFunction("mpi", name, [&](auto*, int, int, auto) {
        MPI_Init(nullptr, nullptr);
        int rank = -1, size = -1;
        MPI_Comm_rank(MPI_COMM_WORLD, &rank);
        MPI_Comm_size(MPI_COMM_WORLD, &size);
        int rc = body(rank, size);
        MPI_Finalize();
        return rc;
    });
    auto req = faabric::util::batchExecFactory("mpi", name, 1);
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

// Windows in the symmetric heap: puts and gets are peer copies
int deviceRmaBody(int rank, int size)
{
    const int n = 4096;
    long* window = nullptr;
    CHECK_RANK(MPI_Alloc_mem(n * sizeof(long), MPI_INFO_FAABRIC_DEVICE, &window) == MPI_SUCCESS);
    std::vector<long> init(n, -1);
    cudaMemcpy(window, init.data(), n * sizeof(long), cudaMemcpyHostToDevice);
    MPI_Win win = nullptr;
    MPI_Win_create(window, n * sizeof(long), sizeof(long), MPI_INFO_NULL, MPI_COMM_WORLD, &win);
    MPI_Win_fence(0, win);
    // host origin -> device window of every rank
    long mine = 100 + rank;
    for (int t = 0; t < size; t++) {
        MPI_Put(&mine, 1, MPI_LONG, t, rank, 1, MPI_LONG, win);
    }
    MPI_Win_fence(0, win);
    std::vector<long> seen(n);
    cudaMemcpy(seen.data(), window, n * sizeof(long), cudaMemcpyDeviceToHost);
    for (int r = 0; r < size; r++) {
        CHECK_RANK(seen[r] == 100 + r);
    }
    CHECK_RANK(seen[size] == -1);
    // device origin <- device window of the right neighbour
    long* strip = nullptr;
    cudaMalloc(&strip, 8 * sizeof(long));
    MPI_Get(strip, 8, MPI_LONG, (rank + 1) % size, 0, 8, MPI_LONG, win);
    MPI_Win_fence(0, win);
    std::vector<long> got(8);
    cudaMemcpy(got.data(), strip, 8 * sizeof(long), cudaMemcpyDeviceToHost);
    CHECK_RANK(got[0] == 100);
    cudaFree(strip);
    MPI_Win_free(&win);
    MPI_Free_mem(window);
    return 0;
}

// Sub-communicator collectives on device buffers go through the p2p layer
int deviceSubCommBody(int rank, int size)
{
    MPI_Comm half = nullptr;
    MPI_Comm_split(MPI_COMM_WORLD, rank % 2, rank, &half);
    int hSize = -1;
    MPI_Comm_size(half, &hSize);
    const int n = 10000;
    int *send = nullptr, *recv = nullptr;
    cudaMalloc(&send, n * sizeof(int));
    cudaMalloc(&recv, n * sizeof(int));
    std::vector<int> host(n, rank + 1);
    cudaMemcpy(send, host.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    MPI_Allreduce(send, recv, n, MPI_INT, MPI_SUM, half);
    cudaMemcpy(host.data(), recv, n * sizeof(int), cudaMemcpyDeviceToHost);
    int expected = 0;
    for (int r = rank % 2; r < size; r += 2) {
        expected += r + 1;
    }
    CHECK_RANK(host[0] == expected && host[n - 1] == expected);
    MPI_Bcast(recv, n, MPI_INT, 0, half);
    cudaFree(send);
    cudaFree(recv);
    MPI_Comm_free(&half);
    return 0;
}
}

TEST_CASE("next: one-sided windows in device memory", "[unverified][mpi]")
{
    NEED_GPU();
    runMpiOnGpus("device-rma", std::max(2, faabric::device::cudaDeviceCountSafe()), deviceRmaBody);
}

TEST_CASE("next: sub-communicator collectives on device buffers", "[unverified][mpi]")
{
    NEED_GPU();
    runMpiOnGpus("device-subcomm", 4, deviceSubCommBody);
}

TEST_CASE("next: device snapshots spill to host, files and back", "[unverified][snapshot]")
{
    NEED_GPU();
    using namespace faabric::util;
    const size_t size = 3 * HOST_PAGE_SIZE + 77;
    std::vector<uint8_t> bytes(size);
    for (size_t i = 0; i < size; i++) {
        bytes[i] = (uint8_t)(i * 13);
    }
    faabric::snapshot::DeviceSnapshot dev(size, 0);
    dev.copyInData(bytes);
    dev.addMergeRegion(128, 8, SnapshotDataType::Long, SnapshotMergeOperation::Sum);
    auto host = dev.spillToHost();
    REQUIRE(host->getDataCopy() == bytes);
    REQUIRE_EQ(host->getMergeRegions().size(), (size_t)1);
    auto back = faabric::snapshot::DeviceSnapshot::fromHost(*host, 0);
    REQUIRE(back->getDataCopy() == bytes);
    REQUIRE_EQ(back->getMergeRegions().size(), (size_t)1);
    const std::string dir = "/tmp/fb_dev_ckpt_" + std::to_string(getpid());
    std::filesystem::create_directories(dir);
    dev.writeToFile(dir + "/d.snap");
    auto fromFile = faabric::snapshot::DeviceSnapshot::readFromFile(dir + "/d.snap", 0);
    REQUIRE(fromFile->getDataCopy() == bytes);
    // registry: device images return to the device
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    reg.clear();
    reg.registerDeviceSnapshot("dev/key", back);
    REQUIRE_EQ(reg.checkpointToDir(dir + "/reg"), (size_t)1);
    reg.clear();
    REQUIRE_EQ(reg.restoreFromDir(dir + "/reg", 0), (size_t)1);
    REQUIRE(reg.deviceSnapshotExists("dev/key"));
    REQUIRE(reg.getDeviceSnapshot("dev/key")->getDataCopy() == bytes);
    reg.clear();
    std::filesystem::remove_all(dir);
}

TEST_CASE("next: a tuning file steers the all-reduce algorithm", "[unverified][device]")
{
    NEED_GPU();
    using namespace faabric::device;
    const std::string path = "/tmp/fb_tuning_gpu_" + std::to_string(getpid()) + ".txt";
    CommTuning t = CommTuning::parse("allreduce 1048576 twoshot\nallreduce 18446744073709551615 oneshot\n");
    faabric::util::writeBytesToFile(path, faabric::util::stringToBytes(t.serialise()));
    setenv("FAABRIC_TUNING_FILE", path.c_str(), 1);
    int nDev = cudaDeviceCountSafe();
    int n = std::max(2, std::min(nDev, 4));
    std::vector<int> devices(n);
    for (int i = 0; i < n; i++) {
        devices[i] = i % nDev;
    }
    CommConfig cfg;
    cfg.heapBytes = 64 << 20;
    auto comms = Communicator::createLocal(n, devices, cfg);
    unsetenv("FAABRIC_TUNING_FILE");
    ::unlink(path.c_str());
    // 4 KiB would be LL by the built-in thresholds; the file says two-shot
    REQUIRE_EQ(comms[0]->pickAllReduceAlgo(4096, false), (int)FB_ALGO_TWOSHOT);
    REQUIRE_EQ(comms[0]->pickAllReduceAlgo(8 << 20, false), (int)FB_ALGO_ONESHOT);
}
