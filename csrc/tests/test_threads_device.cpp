// THREADS fork-join across the per-GPU virtual hosts of one worker:
//  * host memory (CPU): a batch that spans two virtual hosts restores the
//    remote executor from the main thread snapshot and merges its diffs back
//    (reference flow: src/executor/Executor.cpp:111-213,684-730,
//    src/snapshot/SnapshotClient.cpp:76-171, SnapshotServer.cpp:104-142);
//  * device memory ([gpu]): the same batch on DeviceExecutors - restore is a
//    device copy, the merge is ONE fused diff+push kernel per host and only
//    control descriptors cross the RPC layer.
#include "fixtures.h"

#include <faabric/device/cuda_driver.h>
#include <faabric/snapshot/DeviceSnapshot.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/common.h>

#include <cuda_runtime.h>

#include <sys/wait.h>
#include <unistd.h>

using namespace tests;
using faabric::util::SnapshotDataType;
using faabric::util::SnapshotMergeOperation;

TEST_CASE("snapshots: a device image is pushed as a control descriptor, not as bytes", "[snapshot]")
{
    ClusterFixture f(2);
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    faabric::snapshot::DeviceSnapshotDescriptor d;
    d.size = (uint64_t)3 << 30; // 3 GiB "image": nothing of that size crosses the wire
    d.device = 5;
    d.ownerPid = 4242;
    d.devicePtr = 0x7f0000001000ull;
    d.ipcHandle = std::string(64, '\x5a');
    d.mergeRegions.emplace_back(64, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    faabric::snapshot::getSnapshotClient(f.conf.endpointHost)->pushDeviceSnapshot("dev-key", d);
    REQUIRE(reg.deviceDescriptorExists("dev-key"));
    auto got = reg.getDeviceDescriptor("dev-key");
    REQUIRE_EQ(got.size, d.size);
    REQUIRE_EQ(got.device, 5);
    REQUIRE_EQ(got.ownerPid, 4242);
    REQUIRE_EQ(got.devicePtr, d.devicePtr);
    REQUIRE(got.ipcHandle == d.ipcHandle);
    REQUIRE_EQ(got.mergeRegions.size(), 1u);
    REQUIRE_EQ(got.mergeRegions[0].offset, 64u);
    REQUIRE(got.mergeRegions[0].operation == SnapshotMergeOperation::Sum);
    // a host image under the same key is a different thing
    REQUIRE(!reg.snapshotExists("dev-key"));
    reg.deleteDeviceSnapshot("dev-key");
    REQUIRE(!reg.deviceDescriptorExists("dev-key"));

    // mock mode records the descriptor and the device thread result
    faabric::util::setMockMode(true);
    faabric::snapshot::clearMockSnapshotRequests();
    faabric::snapshot::getSnapshotClient("other-host")->pushDeviceSnapshot("k2", d);
    faabric::snapshot::getSnapshotClient("other-host")->pushDeviceThreadResult(7, 99, 3, "k2", 1234);
    auto pushes = faabric::snapshot::getDeviceSnapshotPushes();
    REQUIRE_EQ(pushes.size(), 1u);
    REQUIRE_EQ(std::get<0>(pushes[0]), std::string("other-host"));
    REQUIRE_EQ(std::get<1>(pushes[0]), std::string("k2"));
    REQUIRE_EQ(std::get<2>(pushes[0]).size, d.size);
    auto results = faabric::snapshot::getThreadResults();
    REQUIRE_EQ(results.size(), 1u);
    REQUIRE_EQ(std::get<0>(results[0].second), 99);
    REQUIRE_EQ(std::get<3>(results[0].second), 0); // no diffs travel
    faabric::snapshot::clearMockSnapshotRequests();
    faabric::util::setMockMode(false);
}

TEST_CASE("threads: a batch spanning two virtual hosts of one worker restores and merges", "[executor][threads]")
{
    // this host cannot run anything; two virtual hosts with 2 slots each
    ClusterFixture f(0, 2, 2);
    const int nThreads = 3;
    std::atomic<int> remoteThreads{ 0 };
    int restoresBefore = TestExecutor::restoreCount.load();
    registerTestFunction("demo", "spanning", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto mem = exec->getMemoryView();
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            int t = m.appidx();
            if (m.executedhost() != m.mainhost()) {
                remoteThreads++;
            }
            mem[1024 + t] = (uint8_t)(10 + t);
            __atomic_fetch_add((int*)(mem.data() + 64), t + 1, __ATOMIC_RELAXED);
            return t;
        }
        *(int*)(mem.data() + 64) = 100;
        auto threads = faabric::util::batchExecFactory("demo", "spanning", nThreads);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < nThreads; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
        }
        std::vector<faabric::util::SnapshotMergeRegion> regions = {
            { 64, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum }
        };
        auto results = exec->executeThreads(threads, regions);
        if ((int)results.size() != nThreads) {
            return 1;
        }
        // the main memory was refreshed from the merged snapshot
        mem = exec->getMemoryView();
        int sum = *(int*)(mem.data() + 64);
        m.set_outputdata(std::to_string(sum) + ":" + std::to_string(mem[1025]) + "," + std::to_string(mem[1026]) + "," +
                         std::to_string(mem[1027]));
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "spanning", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0), 20000);
    REQUIRE_EQ(res.returnvalue(), 0);
    // 100 + (2 + 3 + 4), every thread's private byte made it back
    REQUIRE_EQ(res.outputdata(), std::string("109:11,12,13"));
    // one slot was left on the main virtual host: two threads ran elsewhere,
    // in an executor of their own that was restored from the snapshot
    REQUIRE_EQ(remoteThreads.load(), 2);
    REQUIRE(TestExecutor::restoreCount.load() > restoresBefore);
    REQUIRE_EQ(f.sch.getFunctionExecutorCount(req->messages(0)), 2);
    f.awaitBatch(req);
}

// ---------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------
namespace {
class DeviceTestExecutor : public faabric::executor::DeviceExecutor
{
  public:
    explicit DeviceTestExecutor(faabric::Message& msg)
      : DeviceExecutor(msg, (size_t)1 << 20, (size_t)4 << 20)
    {}

    int32_t executeTask(int threadPoolIdx, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req) override
    {
        faabric::Message& msg = *req->mutable_messages(msgIdx);
        auto it = functionTable().find(msg.user() + "/" + msg.function());
        if (it == functionTable().end()) {
            return 0;
        }
        return it->second(this, threadPoolIdx, msgIdx, req);
    }
};

class DeviceTestFactory : public faabric::executor::ExecutorFactory
{
  public:
    std::shared_ptr<faabric::executor::Executor> createExecutor(faabric::Message& msg) override
    {
        return std::make_shared<DeviceTestExecutor>(msg);
    }
};

template<typename T>
T devRead(const uint8_t* p)
{
    T v;
    cudaMemcpy(&v, p, sizeof(T), cudaMemcpyDeviceToHost);
    return v;
}

template<typename T>
void devWrite(uint8_t* p, T v)
{
    cudaMemcpy(p, &v, sizeof(T), cudaMemcpyHostToDevice);
}
}

TEST_CASE("threads on device memory: restore is a device copy, the merge one fused kernel per host", "[gpu][threads]")
{
    if (!faabric::device::cudaAvailable()) {
        SKIP_TEST("no CUDA device");
    }
    ClusterFixture f(0, 2, 2);
    faabric::executor::setExecutorFactory(std::make_shared<DeviceTestFactory>());
    const int nThreads = 3;
    std::atomic<int> remoteThreads{ 0 };
    std::atomic<uint64_t> mergesSeen{ 0 };
    const uint64_t launchesBefore = faabric::snapshot::DeviceSnapshot::getGlobalDiffPushCount();
    registerTestFunction("demo", "devthreads", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto dv = exec->getDeviceMemoryView();
        if (dv.empty()) {
            return 9;
        }
        cudaSetDevice(dv.device);
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            int t = m.appidx();
            if (m.executedhost() != m.mainhost()) {
                remoteThreads++;
            }
            // a private byte, a private far page, and a contribution to the
            // shared Sum word (threads of one host share the memory: serialise)
            devWrite<uint8_t>(dv.ptr + 1024 + t, (uint8_t)(10 + t));
            devWrite<uint32_t>(dv.ptr + 4096 * (10 + t), 0xabc00000u + t);
            static std::mutex sumMx;
            std::lock_guard<std::mutex> lk(sumMx);
            devWrite<int>(dv.ptr + 64, devRead<int>(dv.ptr + 64) + t + 1);
            return t;
        }
        devWrite<int>(dv.ptr + 64, 100);
        auto threads = faabric::util::batchExecFactory("demo", "devthreads", nThreads);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < nThreads; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
        }
        std::vector<faabric::util::SnapshotMergeRegion> regions = {
            { 64, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum }
        };
        auto results = exec->executeThreads(threads, regions);
        if ((int)results.size() != nThreads) {
            return 1;
        }
        mergesSeen = exec->getDeviceMergeCount();
        int sum = devRead<int>(dv.ptr + 64);
        std::string out = std::to_string(sum);
        for (int t = 1; t <= nThreads; t++) {
            out += ":" + std::to_string(devRead<uint8_t>(dv.ptr + 1024 + t));
            if (devRead<uint32_t>(dv.ptr + 4096 * (10 + t)) != 0xabc00000u + t) {
                return 2;
            }
        }
        m.set_outputdata(out);
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "devthreads", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0), 30000);
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::string("109:11:12:13"));
    REQUIRE_EQ(remoteThreads.load(), 2);
    // the main host merged its own thread on the device, the other host its two
    REQUIRE_EQ(mergesSeen.load(), 1u);
    REQUIRE_EQ(faabric::snapshot::DeviceSnapshot::getGlobalDiffPushCount() - launchesBefore, 2u);
    // no host image was ever created for the app
    std::string key = faabric::util::getMainThreadSnapshotKey(req->messages(0));
    REQUIRE(!faabric::snapshot::getSnapshotRegistry().snapshotExists(key));
    REQUIRE(faabric::snapshot::getSnapshotRegistry().deviceSnapshotExists(key));
    f.awaitBatch(req);
}

TEST_CASE("threads on device memory: later fork-joins move only the pages stamped since the last one", "[gpu][threads]")
{
    if (!faabric::device::cudaAvailable()) {
        SKIP_TEST("no CUDA device");
    }
    ClusterFixture f(0, 2, 2);
    faabric::executor::setExecutorFactory(std::make_shared<DeviceTestFactory>());
    const int nThreads = 3;
    const int nRounds = 4;
    std::atomic<int> staleReads{ 0 };
    std::atomic<int> remoteThreads{ 0 };
    registerTestFunction("demo", "devrounds", [&](auto* exec, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        auto dv = exec->getDeviceMemoryView();
        if (dv.empty()) {
            return 9;
        }
        cudaSetDevice(dv.device);
        if (req->type() == faabric::BatchExecuteRequest::THREADS) {
            int t = m.appidx();
            int round = atoi(m.inputdata().c_str());
            if (m.executedhost() != m.mainhost()) {
                remoteThreads++;
            }
            // what the main thread wrote just before this fork must be here,
            // and so must what the OTHER hosts merged in the previous round
            if (devRead<uint32_t>(dv.ptr + 4096 * 20) != (uint32_t)(round + 1) * 1000u) {
                staleReads++;
            }
            if (round > 0) {
                for (int o = 1; o <= nThreads; o++) {
                    // (a thread sharing this host's memory may already have
                    // written this round's value)
                    uint32_t v = devRead<uint32_t>(dv.ptr + 4096 * (10 + o));
                    if (v != (uint32_t)(round * 1000 + o) && v != (uint32_t)((round + 1) * 1000 + o)) {
                        staleReads++;
                    }
                }
            }
            devWrite<uint8_t>(dv.ptr + 1024 + t, (uint8_t)(round * 16 + t));
            devWrite<uint32_t>(dv.ptr + 4096 * (10 + t), (uint32_t)((round + 1) * 1000 + t));
            static std::mutex sumMx;
            std::lock_guard<std::mutex> lk(sumMx);
            devWrite<int>(dv.ptr + 64, devRead<int>(dv.ptr + 64) + t + 1);
            return t;
        }
        devWrite<int>(dv.ptr + 64, 100);
        std::vector<faabric::util::SnapshotMergeRegion> regions = {
            { 64, sizeof(int), SnapshotDataType::Int, SnapshotMergeOperation::Sum }
        };
        for (int round = 0; round < nRounds; round++) {
            // the main thread changes one page between joins
            devWrite<uint32_t>(dv.ptr + 4096 * 20, (uint32_t)(round + 1) * 1000u);
            auto threads = faabric::util::batchExecFactory("demo", "devrounds", nThreads);
            faabric::util::updateBatchExecAppId(threads, m.appid());
            for (int i = 0; i < nThreads; i++) {
                threads->mutable_messages(i)->set_appidx(i + 1);
                threads->mutable_messages(i)->set_groupidx(i + 1);
                threads->mutable_messages(i)->set_inputdata(std::to_string(round));
            }
            auto results = exec->executeThreads(threads, regions);
            if ((int)results.size() != nThreads) {
                return 1;
            }
            // the Sum word accumulates over the rounds, the private bytes and
            // far pages hold this round's values
            if (devRead<int>(dv.ptr + 64) != 100 + (round + 1) * 9) {
                return 2;
            }
            for (int t = 1; t <= nThreads; t++) {
                if (devRead<uint8_t>(dv.ptr + 1024 + t) != (uint8_t)(round * 16 + t) ||
                    devRead<uint32_t>(dv.ptr + 4096 * (10 + t)) != (uint32_t)((round + 1) * 1000 + t)) {
                    return 3;
                }
            }
        }
        m.set_outputdata(std::to_string(devRead<int>(dv.ptr + 64)));
        return 0;
    });
    auto req = faabric::util::batchExecFactory("demo", "devrounds", 1);
    f.plannerCli.callFunctions(req);
    auto res = f.awaitResult(req->messages(0), 60000);
    REQUIRE_EQ(res.returnvalue(), 0);
    REQUIRE_EQ(res.outputdata(), std::to_string(100 + nRounds * 9));
    REQUIRE_EQ(staleReads.load(), 0);
    REQUIRE_EQ(remoteThreads.load(), 2 * nRounds);
    // after the first fork the image keeps page stamps that moved on with every fork and join
    std::string key = faabric::util::getMainThreadSnapshotKey(req->messages(0));
    auto snap = faabric::snapshot::getSnapshotRegistry().getDeviceSnapshot(key);
    REQUIRE(snap->pageStamps() != nullptr);
    REQUIRE_EQ(snap->currentForkStamp(), (uint32_t)(2 * nRounds));
    std::vector<uint32_t> stamps(32);
    cudaMemcpy(stamps.data(), snap->pageStamps(), stamps.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost);
    REQUIRE_EQ(stamps[20], (uint32_t)(2 * nRounds));         // the main thread's page: last fork
    REQUIRE_EQ(stamps[11], (uint32_t)(2 * nRounds + 1));     // a thread's page: last join
    REQUIRE_EQ(stamps[0], (uint32_t)(2 * nRounds + 1));      // the Sum word and the private bytes
    REQUIRE_EQ(stamps[5], 0u);                               // never touched
    f.awaitBatch(req);
}

// Child side of the test below (run as `faabric_tests --ipc-map-child <hex> <size>`):
// maps the parent's image through the descriptor, checks a byte, leaves a mark
int ipcMapChildMain(const char* hexHandle, const char* sizeStr)
{
    faabric::snapshot::DeviceSnapshotDescriptor d;
    std::string hex(hexHandle);
    for (size_t i = 0; i + 1 < hex.size(); i += 2) {
        d.ipcHandle.push_back((char)std::stoi(hex.substr(i, 2), nullptr, 16));
    }
    d.size = std::stoull(sizeStr);
    d.ownerPid = (int)getppid();
    try {
        cudaSetDevice(0);
        auto snap = faabric::snapshot::DeviceSnapshot::fromDescriptor(d);
        auto bytes = snap->getDataCopy(0, 16);
        if (bytes[5] != 77) {
            return 2;
        }
        std::vector<uint8_t> mark = { 123 };
        snap->copyInData(mark, 9);
        return 0;
    } catch (const std::exception& e) {
        fprintf(stderr, "ipc child: %s\n", e.what());
        return 3;
    }
}

TEST_CASE("device snapshot descriptors map the image of another process over CUDA IPC", "[gpu][snapshot]")
{
    if (!faabric::device::cudaAvailable()) {
        SKIP_TEST("no CUDA device");
    }
    auto snap = std::make_shared<faabric::snapshot::DeviceSnapshot>((size_t)1 << 20, 0);
    std::vector<uint8_t> init(16, 0);
    init[5] = 77;
    snap->copyInData(init, 0);
    auto desc = snap->describe();
    REQUIRE_EQ(desc.ipcHandle.size(), sizeof(cudaIpcMemHandle_t));
    REQUIRE_EQ(desc.ownerPid, (int)getpid());
    static const char* digits = "0123456789abcdef";
    std::string hex;
    for (unsigned char c : desc.ipcHandle) {
        hex.push_back(digits[c >> 4]);
        hex.push_back(digits[c & 15]);
    }
    char self[4096];
    ssize_t n = readlink("/proc/self/exe", self, sizeof(self) - 1);
    REQUIRE(n > 0);
    self[n] = 0;
    std::string cmd = std::string(self) + " --ipc-map-child " + hex + " " + std::to_string(desc.size);
    int rc = system(cmd.c_str());
    REQUIRE(WIFEXITED(rc));
    REQUIRE_EQ(WEXITSTATUS(rc), 0);
    // the other process wrote straight into our HBM
    REQUIRE_EQ((int)snap->getDataCopy(9, 1)[0], 123);
}
