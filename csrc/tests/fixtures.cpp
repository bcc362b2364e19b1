#include "fixtures.h"

#include <faabric/transport/common.h>
#include <faabric/util/clock.h>
#include <faabric/util/memory.h>

#include <thread>

namespace tests {

std::map<std::string, TestFunction>& functionTable()
{
    static std::map<std::string, TestFunction> t;
    return t;
}

void registerTestFunction(const std::string& user, const std::string& function, TestFunction fn)
{
    functionTable()[user + "/" + function] = std::move(fn);
}

TestExecutor::TestExecutor(faabric::Message& msg)
  : Executor(msg)
{
    setMemorySize(16 * faabric::util::HOST_PAGE_SIZE);
}

int32_t TestExecutor::executeTask(int threadPoolIdx, int msgIdx, std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    faabric::Message& msg = *req->mutable_messages(msgIdx);
    auto it = functionTable().find(msg.user() + "/" + msg.function());
    if (it == functionTable().end()) {
        // Default body: echo
        msg.set_outputdata(msg.inputdata().empty() ? "ran " + msg.function() : msg.inputdata());
        return 0;
    }
    return it->second(this, threadPoolIdx, msgIdx, req);
}

std::span<uint8_t> TestExecutor::getMemoryView()
{
    return { memory.get(), memorySize };
}

void TestExecutor::setMemorySize(size_t newSize)
{
    if (memory == nullptr) {
        memory = faabric::util::allocateVirtualMemory(MAX_MEMORY);
    }
    if (newSize > memorySize) {
        faabric::util::claimVirtualMemory({ memory.get() + memorySize, newSize - memorySize });
    }
    memorySize = newSize;
}

size_t TestExecutor::getMaxMemorySize()
{
    return MAX_MEMORY;
}

std::atomic<int> TestExecutor::resetCount{ 0 };
std::atomic<int> TestExecutor::restoreCount{ 0 };

void TestExecutor::restore(const std::string& snapshotKey)
{
    restoreCount++;
    auto snap = reg.getSnapshot(snapshotKey);
    setMemorySize(snap->getSize());
    snap->mapToMemory({ memory.get(), snap->getSize() });
}

void TestExecutor::reset(faabric::Message& msg)
{
    resetCount++;
    Executor::reset(msg);
}

std::shared_ptr<faabric::executor::Executor> TestExecutorFactory::createExecutor(faabric::Message& msg)
{
    return std::make_shared<TestExecutor>(msg);
}

ClusterFixture::ClusterFixture(int slots, int nVirtualHosts, int slotsPerVirtualHost)
  : conf(faabric::util::getSystemConfig())
  , planner(faabric::planner::getPlanner())
  , plannerCli(faabric::planner::getPlannerClient())
  , sch(faabric::scheduler::getScheduler())
  , factory(std::make_shared<TestExecutorFactory>())
  , stateServer(faabric::state::getGlobalState())
{
    faabric::util::setMockMode(false);
    conf.reset();
    faabric::executor::setExecutorFactory(factory);
    plannerServer.start();
    planner.reset();
    functionServer.start();
    snapshotServer.start();
    ptpServer.start();
    stateServer.start();
    sch.reset();
    plannerCli.clearCache();

    faabric::HostResources res;
    res.set_slots(slots);
    sch.setThisHostResources(res);
    sch.addHostToGlobalSet();
    for (int i = 0; i < nVirtualHosts; i++) {
        std::string name = "gpu" + std::to_string(i);
        faabric::transport::registerHostAlias(name, conf.endpointHost);
        auto vres = std::make_shared<faabric::HostResources>();
        vres->set_slots(slotsPerVirtualHost);
        sch.addHostToGlobalSet(name, vres);
        virtualHosts.push_back(name);
    }
}

ClusterFixture::~ClusterFixture()
{
    sch.shutdown();
    stateServer.stop();
    ptpServer.stop();
    snapshotServer.stop();
    functionServer.stop();
    planner.reset();
    plannerServer.stop();
    faabric::transport::clearHostAliases();
    faabric::transport::getPointToPointBroker().clear();
    faabric::snapshot::getSnapshotRegistry().clear();
    plannerCli.clearCache();
    conf.reset();
    // The scheduler singleton must be usable by the next fixture
    sch.reset();
}

faabric::Message ClusterFixture::awaitResult(const faabric::Message& msg, int timeoutMs)
{
    return plannerCli.getMessageResult(msg, timeoutMs);
}

std::shared_ptr<faabric::BatchExecuteRequestStatus> ClusterFixture::awaitBatch(
  std::shared_ptr<faabric::BatchExecuteRequest> req,
  int timeoutMs)
{
    auto& clock = faabric::util::getGlobalClock();
    auto t0 = clock.now();
    for (;;) {
        auto status = plannerCli.getBatchResults(req);
        if (status != nullptr && status->finished()) {
            return status;
        }
        if (clock.timeDiff(clock.now(), t0) > timeoutMs) {
            fbtest::fail(__FILE__, __LINE__, "Timed out waiting for batch to finish");
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}

}
