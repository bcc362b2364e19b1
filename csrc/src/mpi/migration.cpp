#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/executor/Executor.h>
#include <faabric/executor/ExecutorContext.h>
#include <faabric/mpi/MpiWorld.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/mpi/migration.h>
#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/scheduler/Scheduler.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/transport/common.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/exception.h>
#include <faabric/util/func.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>

namespace faabric::mpi {

static std::string plannerAddress()
{
    // Same resolution as the planner client: name[:offset]
    const auto& conf = faabric::util::getSystemConfig();
    std::string name = conf.plannerHost;
    std::string suffix;
    size_t colon = name.rfind(':');
    if (colon != std::string::npos && name.find_first_not_of("0123456789", colon + 1) == std::string::npos &&
        colon + 1 < name.size()) {
        suffix = name.substr(colon);
        name = name.substr(0, colon);
    }
    std::string ip = faabric::util::getIPFromHostname(name);
    if (ip.empty()) {
        ip = conf.endpointHost;
    }
    return ip + suffix;
}

// Snapshot of the executor's memory registered under the migration key
static std::string snapshotForMigration(faabric::executor::Executor* exec, int msgId)
{
    std::string key = "migration_" + std::to_string(msgId);
    auto view = exec->getMemoryView();
    auto snap = std::make_shared<faabric::util::SnapshotData>(std::span<const uint8_t>(view.data(), view.size()));
    faabric::snapshot::getSnapshotRegistry().registerSnapshot(key, snap);
    return key;
}

void mpiMigrationPoint(int entrypointArg)
{
    auto ctx = faabric::executor::ExecutorContext::get();
    faabric::Message* call = &ctx->getMsg();
    auto* exec = ctx->getExecutor();
    auto& sch = faabric::scheduler::getScheduler();

    auto migration = sch.checkForMigrationOpportunities(*call);
    const bool appMustFreeze = migration != nullptr && migration->appid() == MUST_FREEZE;
    const std::string arg = std::to_string(entrypointArg);
    auto& reg = faabric::snapshot::getSnapshotRegistry();

    if (appMustFreeze) {
        // Checkpoint to the planner; it re-dispatches when capacity returns
        std::string key = snapshotForMigration(exec, call->id());
        call->set_inputdata(arg);
        call->set_snapshotkey(key);
        faabric::snapshot::getSnapshotClient(plannerAddress())->pushSnapshot(key, reg.getSnapshot(key));
        SPDLOG_INFO("{}:{}:{} Freezing message!", call->appid(), call->groupid(), call->groupidx());
        throw faabric::util::FunctionFrozenException("Freezing MPI rank");
    }

    const bool appMustMigrate = migration != nullptr;
    bool funcMustMigrate = false;
    std::string destination;
    if (appMustMigrate) {
        funcMustMigrate = migration->srchost() != migration->dsthost();
        destination = migration->dsthost();
        // A migration yields a new distribution, hence a new PTP group
        call->set_groupid(migration->groupid());
        if (call->ismpi()) {
            auto& world = getMpiWorldRegistry().getWorld(call->mpiworldid());
            world.prepareMigration(call->groupid(), call->mpirank(), funcMustMigrate);
        }
    }

    if (funcMustMigrate) {
        auto req = faabric::util::batchExecFactory(call->user(), call->function(), 1);
        req->set_type(faabric::BatchExecuteRequest::MIGRATION);
        faabric::util::updateBatchExecAppId(req, migration->appid());
        faabric::util::updateBatchExecGroupId(req, migration->groupid());
        faabric::Message& msg = *req->mutable_messages(0);
        msg.set_inputdata(arg);
        // Same identity on the other side
        msg.set_id(call->id());
        msg.set_appidx(call->appidx());
        msg.set_groupidx(call->groupidx());
        msg.set_mainhost(call->mainhost());
        if (call->ismpi()) {
            msg.set_ismpi(true);
            msg.set_mpiworldid(call->mpiworldid());
            msg.set_mpiworldsize(call->mpiworldsize());
            msg.set_mpirank(call->mpirank());
        }
        msg.set_recordexecgraph(call->recordexecgraph());

        // Only the app's main host pushes snapshots as part of chaining, and
        // we are probably not it: push ours by hand
        std::string key = snapshotForMigration(exec, call->id());
        msg.set_snapshotkey(key);
        if (!faabric::transport::isLocalAddress(faabric::transport::parseHostAddress(destination).ip) ||
            faabric::transport::parseHostAddress(destination).portOffset != faabric::util::getSystemConfig().portOffset) {
            faabric::snapshot::getSnapshotClient(destination)->pushSnapshot(key, reg.getSnapshot(key));
        }

        SPDLOG_DEBUG("Migrating {}:{}:{} from {} to {}", call->appid(), call->groupid(), call->groupidx(), migration->srchost(), destination);
        faabric::scheduler::getFunctionCallClient(destination)->executeFunctions(req);
        if (call->recordexecgraph()) {
            faabric::util::logChainedFunction(*call, msg);
        }
        throw faabric::util::FunctionMigratedException("Migrating MPI rank");
    }

    // Staying, but somebody moved: line up with the new group
    if (appMustMigrate) {
        faabric::transport::getPointToPointBroker().postMigrationHook(call->groupid(), call->groupidx());
    }
}

}
