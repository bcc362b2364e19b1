// BASELINE config 5: batch-schedule N no-op functions over H hosts, fan-out +
// fan-in, end to end.  Prints one JSON line.
//   planner_bench [--functions 1024] [--hosts 8] [--iters 20] [--warmup 3]
#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/LocalCluster.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <numeric>

#include <cxxabi.h>
#include <dlfcn.h>
#include <execinfo.h>
#include <signal.h>
#include <sys/time.h>

// ---- --profile: a tiny sampling profiler (SIGPROF every 200 us of CPU time,
// delivered to whichever thread is running); prints the hottest functions by
// inclusive samples.  No external tools are available in the build image.
namespace {
constexpr int MAX_SAMPLES = 200000;
constexpr int DEPTH = 24;
void* g_samples[MAX_SAMPLES][DEPTH];
int g_depths[MAX_SAMPLES];
std::atomic<int> g_nSamples{ 0 };

void profHandler(int, siginfo_t*, void*)
{
    int i = g_nSamples.fetch_add(1);
    if (i < MAX_SAMPLES) {
        g_depths[i] = backtrace(g_samples[i], DEPTH);
    }
}

void startProfiler()
{
    void* warm[4];
    backtrace(warm, 4); // loads libgcc outside the handler
    struct sigaction sa{};
    sa.sa_sigaction = profHandler;
    sa.sa_flags = SA_SIGINFO | SA_RESTART;
    sigaction(SIGPROF, &sa, nullptr);
    itimerval tv{};
    tv.it_interval.tv_usec = 200;
    tv.it_value.tv_usec = 200;
    setitimer(ITIMER_PROF, &tv, nullptr);
}

void reportProfiler()
{
    itimerval tv{};
    setitimer(ITIMER_PROF, &tv, nullptr);
    int n = std::min(g_nSamples.load(), MAX_SAMPLES);
    std::map<std::string, int> inclusive, self, under;
    // PROFILE_ROOT=<substring>: a third table, inclusive counts of the samples
    // whose stack contains a matching frame
    const char* rootEnv = getenv("PROFILE_ROOT");
    std::string root = rootEnv ? rootEnv : "";
    int nUnder = 0;
    for (int i = 0; i < n; i++) {
        std::map<std::string, bool> seen;
        std::vector<std::string> names;
        for (int d = 2; d < g_depths[i]; d++) { // skip handler + signal frame
            Dl_info info;
            std::string name = "?";
            if (dladdr(g_samples[i][d], &info) && info.dli_sname) {
                int st = 0;
                char* dem = abi::__cxa_demangle(info.dli_sname, nullptr, nullptr, &st);
                name = st == 0 && dem ? dem : info.dli_sname;
                free(dem);
            }
            if (name.size() > 90) {
                name.resize(90);
            }
            if (d == 2) {
                self[name]++;
            }
            if (!seen[name]) {
                seen[name] = true;
                inclusive[name]++;
                names.push_back(name);
            }
        }
        if (!root.empty()) {
            bool has = false;
            for (auto& nm : names) {
                has = has || nm.find(root) != std::string::npos;
            }
            if (has) {
                nUnder++;
                for (auto& nm : names) {
                    under[nm]++;
                }
            }
        }
    }
    auto top = [&](std::map<std::string, int>& m, const char* title) {
        std::vector<std::pair<int, std::string>> v;
        for (auto& [k, c] : m) {
            v.emplace_back(c, k);
        }
        std::sort(v.rbegin(), v.rend());
        fprintf(stderr, "---- %s (of %d samples) ----\n", title, n);
        for (size_t i = 0; i < v.size() && i < 28; i++) {
            fprintf(stderr, "%6.2f%%  %s\n", 100.0 * v[i].first / std::max(n, 1), v[i].second.c_str());
        }
    };
    top(self, "self");
    top(inclusive, "inclusive");
    if (!root.empty()) {
        fprintf(stderr, "(%d samples under *%s*)\n", nUnder, root.c_str());
        top(under, "inclusive, under the root");
    }
}
}

using namespace faabric::executor;

class NoopExecutor : public Executor
{
  public:
    explicit NoopExecutor(faabric::Message& msg)
      : Executor(msg)
    {}

    int32_t executeTask(int, int, std::shared_ptr<faabric::BatchExecuteRequest>) override
    {
        // when did the first / last function of the batch actually run
        int64_t now = std::chrono::steady_clock::now().time_since_epoch().count();
        int64_t seen = firstExec.load(std::memory_order_relaxed);
        while (now < seen && !firstExec.compare_exchange_weak(seen, now)) {
        }
        seen = lastExec.load(std::memory_order_relaxed);
        while (now > seen && !lastExec.compare_exchange_weak(seen, now)) {
        }
        return 0;
    }

    static std::atomic<int64_t> firstExec;
    static std::atomic<int64_t> lastExec;
};

std::atomic<int64_t> NoopExecutor::firstExec{ INT64_MAX };
std::atomic<int64_t> NoopExecutor::lastExec{ 0 };

class NoopFactory : public ExecutorFactory
{
  protected:
    std::shared_ptr<Executor> createExecutor(faabric::Message& msg) override
    {
        return std::make_shared<NoopExecutor>(msg);
    }
};

int main(int argc, char** argv)
{
    int nFunctions = 1024, nHosts = 8, iters = 20, warmup = 3;
    bool profile = false;
    // --mode refcpu: the reference's control-plane design on the same box -
    // every request and result is encoded, sent over a (loopback) socket and
    // decoded again, results funnel through the planner's RPC workers
    // (reference: src/planner/Planner.cpp:807-1394, PlannerServer.cpp:227-248,
    // FunctionCallClient.cpp:66-131).  Default: this repo's typed in-process
    // hand-off between the planner and the per-GPU hosts of one worker.
    std::string mode = "native";
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--profile")) {
            profile = true;
        }
        if (!strcmp(argv[i], "--mode") && i + 1 < argc) {
            mode = argv[i + 1];
        }
    }
    if (mode == "refcpu") {
        setenv("FAABRIC_INPROC_RPC", "0", 1);
    }
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--profile")) {
            i--;
            continue;
        }
        if (!strcmp(argv[i], "--functions")) {
            nFunctions = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--hosts")) {
            nHosts = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--iters")) {
            iters = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--warmup")) {
            warmup = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--mode")) {
            // parsed above
        }
    }
    setenv("LOG_LEVEL", "warn", 0);
    faabric::util::getSystemConfig().reset();
    faabric::util::initLogging();

    int perHost = (nFunctions + nHosts - 1) / nHosts;
    faabric::runner::LocalCluster cluster(std::make_shared<NoopFactory>(), nHosts, perHost);
    auto& cli = faabric::planner::getPlannerClient();

    std::vector<double> totalUs, scheduleUs, firstRunUs, lastRunUs;
    for (int it = 0; it < warmup + iters; it++) {
        if (profile && it == warmup) {
            startProfiler();
        }
        auto req = faabric::util::batchExecFactory("bench", "noop", nFunctions);
        NoopExecutor::firstExec.store(INT64_MAX);
        NoopExecutor::lastExec.store(0);
        auto t0 = std::chrono::steady_clock::now();
        auto decision = cli.callFunctions(req);
        auto t1 = std::chrono::steady_clock::now();
        if (decision.nFunctions != nFunctions) {
            fprintf(stderr, "scheduling failed at iteration %d: appId %d nFunctions %d\n", it, (int)decision.appId, decision.nFunctions);
            return 1;
        }
        auto status = cluster.awaitBatch(req);
        auto t2 = std::chrono::steady_clock::now();
        if (status->messageresults_size() != nFunctions) {
            fprintf(stderr, "missing results\n");
            return 1;
        }
        if (it >= warmup) {
            scheduleUs.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            totalUs.push_back(std::chrono::duration<double, std::micro>(t2 - t0).count());
            int64_t base = t0.time_since_epoch().count();
            firstRunUs.push_back((double)(NoopExecutor::firstExec.load() - base) / 1e3);
            lastRunUs.push_back((double)(NoopExecutor::lastExec.load() - base) / 1e3);
        }
    }
    if (profile) {
        reportProfiler();
    }
    std::sort(totalUs.begin(), totalUs.end());
    std::sort(scheduleUs.begin(), scheduleUs.end());
    std::sort(firstRunUs.begin(), firstRunUs.end());
    std::sort(lastRunUs.begin(), lastRunUs.end());
    double med = totalUs[totalUs.size() / 2];
    printf("{\"bench\": \"planner_fanout\", \"mode\": \"%s\", \"functions\": %d, \"hosts\": %d, \"iters\": %d, "
           "\"e2e_us_median\": %.1f, \"e2e_us_min\": %.1f, \"e2e_us_max\": %.1f, \"schedule_us_median\": %.1f, "
           "\"first_function_runs_at_us\": %.1f, \"last_function_runs_at_us\": %.1f, "
           "\"functions_per_s\": %.0f}\n",
           mode.c_str(),
           nFunctions,
           nHosts,
           iters,
           med,
           totalUs.front(),
           totalUs.back(),
           scheduleUs[scheduleUs.size() / 2],
           firstRunUs[firstRunUs.size() / 2],
           lastRunUs[lastRunUs.size() / 2],
           nFunctions / med * 1e6);
    return 0;
}
