// THREADS fork-join through the runtime (planner -> scheduler -> executors ->
// snapshot registry), end to end: a main function forks one thread per
// (virtual GPU) host, every thread dirties a share of the function memory, the
// join merges every host's changes back into the main image.
//   threads_bench [--memory device|host] [--hosts H] [--mem-mb 1024]
//                 [--dirty-pct 1] [--iters 10] [--warmup 2]
// --memory device: function memory in HBM (DeviceExecutor): restore is a peer
//   copy, the merge one fused diff+push kernel per host, only descriptors
//   cross the RPC layer.
// --memory host: the reference's design on this box - mprotect/SIGSEGV dirty
//   tracking, CoW-mapped restore, byte diffs pushed through the snapshot
//   server (reference: src/executor/Executor.cpp:111-213,684-730,
//   src/snapshot/SnapshotClient.cpp:76-171).
// Prints one JSON line.
#include <faabric/executor/Executor.h>
#include <faabric/executor/ExecutorFactory.h>
#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/LocalCluster.h>
#include <faabric/snapshot/DeviceSnapshot.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/hwloc.h>
#include <faabric/util/logging.h>
#include <faabric/util/memory.h>
#include <faabric/util/snapshot.h>

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>

using namespace faabric::executor;

namespace {
struct Params
{
    bool device = true;
    int hosts = 0;
    size_t memBytes = (size_t)1 << 30;
    double dirtyPct = 1.0;
    int iters = 10;
    int warmup = 2;
} P;

constexpr size_t PAGE = 4096;
std::vector<double> g_iterMs;
std::atomic<int> g_errors{ 0 };
std::atomic<uint64_t> g_dirtyBytes{ 0 };

// Pages dirtied by thread t of nThreads: every `stride`-th page, offset by t
size_t pageStride(int nThreads)
{
    size_t s = (size_t)(100.0 / P.dirtyPct);
    return std::max<size_t>(s, 1) * (size_t)nThreads;
}

uint8_t markOf(int iter, int t)
{
    return (uint8_t)(((iter * 31 + t * 7) & 0x7f) | 0x80);
}

int32_t runThread(Executor* exec, int iter, int t, int nThreads)
{
    const size_t nPages = P.memBytes / PAGE;
    const size_t stride = pageStride(nThreads);
    const size_t first = (size_t)(t - 1) * (stride / nThreads);
    if (first >= nPages) {
        return 0;
    }
    const size_t count = (nPages - first + stride - 1) / stride;
    const uint8_t mark = markOf(iter, t);
    if (P.device) {
        auto dv = exec->getDeviceMemoryView();
        cudaSetDevice(dv.device);
        // one strided fill: `count` pages, one every `stride`
        cudaError_t e = cudaMemset2D(dv.ptr + first * PAGE, stride * PAGE, mark, PAGE, count);
        if (e == cudaSuccess) {
            e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess) {
            fprintf(stderr, "thread %d: %s\n", t, cudaGetErrorString(e));
            return 1;
        }
    } else {
        auto mem = exec->getMemoryView();
        for (size_t p = first; p < nPages; p += stride) {
            memset(mem.data() + p * PAGE, mark, PAGE);
        }
    }
    g_dirtyBytes.fetch_add(count * PAGE);
    return 0;
}

int32_t runMain(Executor* exec, faabric::Message& m)
{
    const int nThreads = P.hosts;
    const size_t nPages = P.memBytes / PAGE;
    const size_t stride = pageStride(nThreads);
    for (int it = 0; it < P.warmup + P.iters; it++) {
        auto threads = faabric::util::batchExecFactory("bench", "fork", nThreads);
        faabric::util::updateBatchExecAppId(threads, m.appid());
        for (int i = 0; i < nThreads; i++) {
            threads->mutable_messages(i)->set_appidx(i + 1);
            threads->mutable_messages(i)->set_groupidx(i + 1);
            threads->mutable_messages(i)->set_inputdata(std::to_string(it));
        }
        auto t0 = std::chrono::steady_clock::now();
        auto results = exec->executeThreads(threads, {});
        auto t1 = std::chrono::steady_clock::now();
        for (auto& r : results) {
            if (r.second != 0) {
                g_errors++;
            }
        }
        // every thread's first and last page made it into the main memory
        for (int t = 1; t <= nThreads; t++) {
            size_t first = (size_t)(t - 1) * (stride / nThreads);
            size_t last = first + ((nPages - first - 1) / stride) * stride;
            for (size_t p : { first, last }) {
                uint8_t got[2] = { 0, 0 };
                if (P.device) {
                    auto dv = exec->getDeviceMemoryView();
                    cudaSetDevice(dv.device);
                    cudaMemcpy(&got[0], dv.ptr + p * PAGE, 1, cudaMemcpyDeviceToHost);
                    cudaMemcpy(&got[1], dv.ptr + p * PAGE + PAGE - 1, 1, cudaMemcpyDeviceToHost);
                } else {
                    auto mem = exec->getMemoryView();
                    got[0] = mem[p * PAGE];
                    got[1] = mem[p * PAGE + PAGE - 1];
                }
                if (got[0] != markOf(it, t) || got[1] != markOf(it, t)) {
                    if (g_errors++ < 4) {
                        fprintf(stderr, "iter %d thread %d page %zu: got %d,%d want %d\n", it, t, p, got[0], got[1], markOf(it, t));
                    }
                }
            }
        }
        if (it >= P.warmup) {
            g_iterMs.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        } else if (it == 0) {
            fprintf(stderr, "first fork-join (cold: executors created, image registered): %.2f ms\n",
                    std::chrono::duration<double, std::milli>(t1 - t0).count());
        }
    }
    return g_errors.load() == 0 ? 0 : 1;
}

int32_t dispatch(Executor* exec, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    faabric::Message& m = *req->mutable_messages(msgIdx);
    if (req->type() == faabric::BatchExecuteRequest::THREADS) {
        return runThread(exec, atoi(m.inputdata().c_str()), m.appidx(), P.hosts);
    }
    return runMain(exec, m);
}

class DeviceBenchExecutor : public DeviceExecutor
{
  public:
    explicit DeviceBenchExecutor(faabric::Message& msg)
      : DeviceExecutor(msg, P.memBytes, P.memBytes)
    {}

    int32_t executeTask(int, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req) override
    {
        return dispatch(this, msgIdx, req);
    }
};

class HostBenchExecutor : public Executor
{
  public:
    explicit HostBenchExecutor(faabric::Message& msg)
      : Executor(msg)
    {
        memory = faabric::util::allocateVirtualMemory(P.memBytes);
        faabric::util::claimVirtualMemory({ memory.get(), P.memBytes });
        memset(memory.get(), 1, P.memBytes); // resident, like the device image
    }

    int32_t executeTask(int, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req) override
    {
        return dispatch(this, msgIdx, req);
    }

    std::span<uint8_t> getMemoryView() override { return { memory.get(), P.memBytes }; }

    void setMemorySize(size_t) override {}

    size_t getMaxMemorySize() override { return P.memBytes; }

  private:
    faabric::util::MemoryRegion memory;
};

class BenchFactory : public ExecutorFactory
{
  protected:
    std::shared_ptr<Executor> createExecutor(faabric::Message& msg) override
    {
        if (P.device) {
            return std::make_shared<DeviceBenchExecutor>(msg);
        }
        return std::make_shared<HostBenchExecutor>(msg);
    }
};
}

int main(int argc, char** argv)
{
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i], v = argv[i + 1];
        if (k == "--memory") {
            P.device = v != "host";
        } else if (k == "--hosts") {
            P.hosts = atoi(v.c_str());
        } else if (k == "--mem-mb") {
            P.memBytes = (size_t)atoll(v.c_str()) << 20;
        } else if (k == "--dirty-pct") {
            P.dirtyPct = atof(v.c_str());
        } else if (k == "--iters") {
            P.iters = atoi(v.c_str());
        } else if (k == "--warmup") {
            P.warmup = atoi(v.c_str());
        }
    }
    setenv("LOG_LEVEL", "warn", 0);
    faabric::util::getSystemConfig().reset();
    faabric::util::initLogging();
    int nGpus = 0;
    if (P.device && (cudaGetDeviceCount(&nGpus) != cudaSuccess || nGpus == 0)) {
        printf("{\"bench\": \"threads_forkjoin\", \"unavailable\": \"no CUDA device\"}\n");
        return 0;
    }
    if (P.hosts <= 0) {
        P.hosts = P.device ? std::max(nGpus, 2) : 2;
    }

    // one slot per virtual host for the threads, one more on the main host for
    // the forking function itself
    faabric::runner::LocalCluster cluster(std::make_shared<BenchFactory>(), P.hosts, 2);
    auto& cli = faabric::planner::getPlannerClient();
    const uint64_t launches0 = faabric::snapshot::DeviceSnapshot::getGlobalDiffPushCount();
    auto req = faabric::util::batchExecFactory("bench", "fork", 1);
    auto tAll0 = std::chrono::steady_clock::now();
    cli.callFunctions(req);
    auto status = cluster.awaitBatch(req, 600000);
    auto tAll1 = std::chrono::steady_clock::now();
    int mainRv = -1;
    for (const auto& r : status->messageresults()) {
        if (r.id() == req->messages(0).id()) {
            mainRv = r.returnvalue();
        }
    }
    if (mainRv != 0 || g_iterMs.empty()) {
        fprintf(stderr, "fork-join bench failed (%d errors)\n", g_errors.load());
        return 1;
    }
    const uint64_t launches = faabric::snapshot::DeviceSnapshot::getGlobalDiffPushCount() - launches0;
    std::sort(g_iterMs.begin(), g_iterMs.end());
    const double med = g_iterMs[g_iterMs.size() / 2];
    const double dirtyPerIter = (double)g_dirtyBytes.load() / (P.warmup + P.iters);
    printf("{\"bench\": \"threads_forkjoin\", \"memory\": \"%s\", \"hosts\": %d, \"gpus\": %d, \"threads\": %d, "
           "\"mem_bytes\": %zu, \"dirty_pct\": %.3f, \"dirty_bytes_per_join\": %.0f, \"iters\": %d, "
           "\"ms_median\": %.3f, \"ms_min\": %.3f, \"ms_max\": %.3f, \"image_gb_per_s\": %.1f, "
           "\"diff_push_kernels\": %llu, \"total_s\": %.2f, \"verified\": true}\n",
           P.device ? "device" : "host",
           P.hosts,
           nGpus,
           P.hosts,
           P.memBytes,
           P.dirtyPct,
           dirtyPerIter,
           P.iters,
           med,
           g_iterMs.front(),
           g_iterMs.back(),
           (double)P.memBytes / 1e9 / (med / 1e3),
           (unsigned long long)launches,
           std::chrono::duration<double>(tAll1 - tAll0).count());
    return 0;
}
