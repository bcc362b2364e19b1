// BASELINE config 5: batch-schedule N no-op functions over H hosts, fan-out +
// fan-in, end to end.  Prints one JSON line.
//   planner_bench [--functions 1024] [--hosts 8] [--iters 20] [--warmup 3]
#include <faabric/planner/PlannerClient.h>
#include <faabric/runner/LocalCluster.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <numeric>

using namespace faabric::executor;

class NoopExecutor : public Executor
{
  public:
    explicit NoopExecutor(faabric::Message& msg)
      : Executor(msg)
    {}

    int32_t executeTask(int, int, std::shared_ptr<faabric::BatchExecuteRequest>) override { return 0; }
};

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
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--functions")) {
            nFunctions = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--hosts")) {
            nHosts = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--iters")) {
            iters = atoi(argv[i + 1]);
        } else if (!strcmp(argv[i], "--warmup")) {
            warmup = atoi(argv[i + 1]);
        }
    }
    setenv("LOG_LEVEL", "warn", 0);
    faabric::util::getSystemConfig().reset();
    faabric::util::initLogging();

    int perHost = (nFunctions + nHosts - 1) / nHosts;
    faabric::runner::LocalCluster cluster(std::make_shared<NoopFactory>(), nHosts, perHost);
    auto& cli = faabric::planner::getPlannerClient();

    std::vector<double> totalUs, scheduleUs;
    for (int it = 0; it < warmup + iters; it++) {
        auto req = faabric::util::batchExecFactory("bench", "noop", nFunctions);
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
        }
    }
    std::sort(totalUs.begin(), totalUs.end());
    std::sort(scheduleUs.begin(), scheduleUs.end());
    double med = totalUs[totalUs.size() / 2];
    printf("{\"bench\": \"planner_fanout\", \"functions\": %d, \"hosts\": %d, \"iters\": %d, "
           "\"e2e_us_median\": %.1f, \"e2e_us_min\": %.1f, \"e2e_us_max\": %.1f, \"schedule_us_median\": %.1f, "
           "\"functions_per_s\": %.0f}\n",
           nFunctions,
           nHosts,
           iters,
           med,
           totalUs.front(),
           totalUs.back(),
           scheduleUs[scheduleUs.size() / 2],
           nFunctions / med * 1e6);
    return 0;
}
