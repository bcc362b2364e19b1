#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/planner/Planner.h>
#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/snapshot/SnapshotClient.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/transport/common.h>
#include <faabric/util/batch.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/environment.h>
#include <faabric/util/func.h>
#include <faabric/util/gids.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>

#include <algorithm>
#include <sstream>

// Group id marking a preloaded decision that already owns its slots/ports
#define FIXED_SIZE_PRELOADED_DECISION_GROUPID -99

namespace faabric::planner {

using faabric::batch_scheduler::DecisionType;
using faabric::batch_scheduler::SchedulingDecision;

// ---------------------------------------------------------------------------
// Host bookkeeping helpers
// ---------------------------------------------------------------------------
static void claimHostSlots(const std::shared_ptr<Host>& host, int n = 1)
{
    host->set_usedslots(host->usedslots() + n);
}

static void releaseHostSlots(const std::shared_ptr<Host>& host, int n = 1)
{
    host->set_usedslots(std::max(0, host->usedslots() - n));
}

// "MPI ports" are per-host slot indices (mailbox / stream slot on a GPU host)
static int claimHostMpiPort(const std::shared_ptr<Host>& host)
{
    for (int i = 0; i < host->mpiports_size(); i++) {
        auto* p = host->mutable_mpiports(i);
        if (!p->used()) {
            p->set_used(true);
            return p->port();
        }
    }
    SPDLOG_ERROR("Ran out of MPI ports on host {}", host->ip());
    throw std::runtime_error("Ran out of MPI ports on host");
}

static void releaseHostMpiPort(const std::shared_ptr<Host>& host, int port)
{
    if (port <= 0) {
        return;
    }
    for (int i = 0; i < host->mpiports_size(); i++) {
        auto* p = host->mutable_mpiports(i);
        if (p->port() == port) {
            p->set_used(false);
            return;
        }
    }
}

static std::string hostStateString(
  const std::map<std::string, std::shared_ptr<Host>>& hostMap)
{
    std::ostringstream os;
    os << "--- Planner host state ---\n";
    for (const auto& [ip, h] : hostMap) {
        os << ip << ": " << h->usedslots() << "/" << h->slots() << "\n";
    }
    return os.str();
}

static faabric::batch_scheduler::HostMap toSchedulerHostMap(
  const std::map<std::string, std::shared_ptr<Host>>& hostMapIn,
  const std::set<std::string>& evicted)
{
    faabric::batch_scheduler::HostMap out;
    for (const auto& [ip, h] : hostMapIn) {
        auto hs = std::make_shared<faabric::batch_scheduler::HostState>(
          h->ip(), h->slots(), h->usedslots());
        if (evicted.count(ip) > 0) {
            hs->ip = MUST_EVICT_IP;
        }
        out[ip] = hs;
    }
    return out;
}

// Free slots an OpenMP app could grow into on its main host
static int availableOpenMpSlots(
  int appId,
  const std::string& mainHost,
  const std::map<std::string, std::shared_ptr<Host>>& hostMap,
  const faabric::batch_scheduler::InFlightReqs& inFlight)
{
    auto it = hostMap.find(mainHost);
    if (it == hostMap.end()) {
        return 0;
    }
    int avail = it->second->slots() - it->second->usedslots();
    // Other OpenMP apps on this host that have announced more threads than
    // they currently run keep their reservation
    for (const auto& [otherId, pair] : inFlight) {
        if (otherId == appId || pair.first->messages_size() == 0 ||
            !pair.first->messages(0).isomp() || pair.second->hosts.empty() ||
            pair.second->hosts.at(0) != mainHost) {
            continue;
        }
        int reserved =
          pair.first->messages(0).ompnumthreads() - pair.first->messages_size();
        if (reserved > 0) {
            avail -= reserved;
        }
    }
    return std::max(0, avail);
}

// ---------------------------------------------------------------------------
// Planner
// ---------------------------------------------------------------------------
Planner::Planner()
  : snapshotRegistry(faabric::snapshot::getSnapshotRegistry())
{
    auto& conf = faabric::util::getSystemConfig();
    config.set_ip(conf.endpointHost);
    config.set_hosttimeout(
      std::stoi(faabric::util::getEnvVar("PLANNER_HOST_KEEPALIVE_TIMEOUT", "5")));
    config.set_numthreadshttpserver(
      std::stoi(faabric::util::getEnvVar("PLANNER_HTTP_SERVER_THREADS", "4")));
    state.policy = conf.batchSchedulerMode;
    printConfig();
}

PlannerConfig Planner::getConfig()
{
    return config;
}

void Planner::printConfig() const
{
    SPDLOG_INFO("--- Planner Conifg ---");
    SPDLOG_INFO("HOST_KEEP_ALIVE_TIMEOUT    {}", config.hosttimeout());
    SPDLOG_INFO("HTTP_SERVER_THREADS        {}", config.numthreadshttpserver());
}

std::string Planner::getPolicy()
{
    std::shared_lock<std::shared_mutex> lock(plannerMx);
    return state.policy;
}

void Planner::setPolicy(const std::string& newPolicy)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    // Throws for an unknown policy and leaves the old one in place
    std::string old = faabric::util::getSystemConfig().batchSchedulerMode;
    try {
        faabric::batch_scheduler::resetBatchScheduler(newPolicy);
    } catch (const std::exception&) {
        faabric::batch_scheduler::resetBatchScheduler(old);
        throw;
    }
    state.policy = newPolicy;
}

bool Planner::reset()
{
    SPDLOG_INFO("Resetting planner");
    flushSchedulingState();
    flushHosts();
    return true;
}

bool Planner::flush(faabric::planner::FlushType flushType)
{
    switch (flushType) {
        case FlushType::Hosts:
            SPDLOG_INFO("Planner flushing available hosts state");
            flushHosts();
            return true;
        case FlushType::Executors:
            SPDLOG_INFO("Planner flushing executors");
            flushExecutors();
            return true;
        case FlushType::SchedulingState:
            SPDLOG_INFO("Planner flushing scheduling state");
            flushSchedulingState();
            return true;
        default:
            SPDLOG_ERROR("Unrecognised flush type");
            return false;
    }
}

void Planner::flushHosts()
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    state.hostMap.clear();
}

void Planner::flushExecutors()
{
    auto hosts = getAvailableHosts();
    for (const auto& h : hosts) {
        SPDLOG_INFO("Planner sending EXECUTOR flush to {}", h->ip());
        faabric::scheduler::getFunctionCallClient(h->ip())->sendFlush();
    }
}

void Planner::flushSchedulingState()
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    state.policy = "bin-pack";
    faabric::batch_scheduler::resetBatchScheduler("bin-pack");
    state.stateMains.clear();
    state.inFlightReqs.clear();
    state.finishedInFlight.clear();
    state.inFlightIdPos.clear();
    appFinishedCv.notify_all();
    state.appResults.clear();
    state.appResultWaiters.clear();
    state.preloadedSchedulingDecisions.clear();
    state.evictedRequests.clear();
    state.nextEvictedHostIps.clear();
    state.numMigrations = 0;
    // Slots of every host become free again
    for (auto& [ip, h] : state.hostMap) {
        h->set_usedslots(0);
        for (int i = 0; i < h->mpiports_size(); i++) {
            h->mutable_mpiports(i)->set_used(false);
        }
    }
}

// ---------------------------------------------------------------------------
// Membership
// ---------------------------------------------------------------------------
void Planner::setHostKeepAliveTimeout(int seconds)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    config.set_hosttimeout(seconds);
}

bool Planner::isHostExpired(std::shared_ptr<Host> host, long epochTimeMs)
{
    if (epochTimeMs == 0) {
        epochTimeMs = faabric::util::getGlobalClock().epochMillis();
    }
    long timeoutMs = (long)config.hosttimeout() * 1000;
    return (epochTimeMs - host->registerts().epochms()) > timeoutMs;
}

std::vector<std::shared_ptr<Host>> Planner::getAvailableHosts()
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    // Lazily drop hosts that stopped sending keep-alives
    long now = faabric::util::getGlobalClock().epochMillis();
    for (auto it = state.hostMap.begin(); it != state.hostMap.end();) {
        if (isHostExpired(it->second, now)) {
            SPDLOG_WARN("Planner removing expired host {}", it->first);
            it = state.hostMap.erase(it);
        } else {
            ++it;
        }
    }
    std::vector<std::shared_ptr<Host>> out;
    out.reserve(state.hostMap.size());
    for (const auto& [ip, h] : state.hostMap) {
        out.push_back(h);
    }
    return out;
}

bool Planner::registerHost(const Host& hostIn, bool overwrite)
{
    SPDLOG_TRACE("Planner received request to register host {}", hostIn.ip());
    if (hostIn.slots() < 0) {
        SPDLOG_ERROR("Received erroneous request to register host {} with {} slots",
                     hostIn.ip(),
                     hostIn.slots());
        return false;
    }
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    auto it = state.hostMap.find(hostIn.ip());
    if (it == state.hostMap.end() || isHostExpired(it->second)) {
        if (it != state.hostMap.end()) {
            state.hostMap.erase(it);
        }
        SPDLOG_INFO("Registering host {} with {} slots", hostIn.ip(), hostIn.slots());
        auto h = std::make_shared<Host>(hostIn);
        // One MPI port (mailbox slot) per execution slot
        h->clear_mpiports();
        for (int i = 0; i < h->slots(); i++) {
            auto* p = h->add_mpiports();
            p->set_port(MPI_BASE_PORT + i);
            p->set_used(false);
        }
        h->mutable_registerts()->set_epochms(
          faabric::util::getGlobalClock().epochMillis());
        state.hostMap.emplace(hostIn.ip(), std::move(h));
    } else {
        if (overwrite) {
            SPDLOG_INFO("Overwriting host {} with {} slots (used {})",
                        hostIn.ip(),
                        hostIn.slots(),
                        hostIn.usedslots());
            it->second->set_slots(hostIn.slots());
            it->second->set_usedslots(hostIn.usedslots());
            it->second->clear_mpiports();
            for (int i = 0; i < hostIn.slots(); i++) {
                auto* p = it->second->add_mpiports();
                p->set_port(MPI_BASE_PORT + i);
                p->set_used(false);
            }
        }
        // Keep-alive
        it->second->mutable_registerts()->set_epochms(
          faabric::util::getGlobalClock().epochMillis());
    }
    return true;
}

void Planner::removeHost(const Host& hostIn)
{
    SPDLOG_DEBUG("Planner received request to remove host {}", hostIn.ip());
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    state.hostMap.erase(hostIn.ip());
}

// ---------------------------------------------------------------------------
// Results
// ---------------------------------------------------------------------------
void Planner::setMessageResult(std::shared_ptr<faabric::Message> msg)
{
    // A migrated message carries on elsewhere: its result comes later
    if (msg->returnvalue() == MIGRATED_FUNCTION_RETURN_VALUE) {
        return;
    }
    std::vector<std::string> toNotify;
    {
        std::unique_lock<std::shared_mutex> lock(plannerMx);
        recordResultLocked(msg, toNotify);
    }
    // Notify outside the lock
    for (const auto& host : toNotify) {
        faabric::scheduler::getFunctionCallClient(host)->setMessageResult(msg);
    }
}

// A fan-in of N results costs ONE exclusive acquisition of the planner's lock
void Planner::setMessageResults(const std::vector<std::shared_ptr<faabric::Message>>& msgs)
{
    std::vector<std::pair<std::shared_ptr<faabric::Message>, std::vector<std::string>>> notify;
    {
        std::unique_lock<std::shared_mutex> lock(plannerMx);
        // results of a fan-in mostly belong to one app and a handful of
        // hosts: the look-ups they share are done once
        ResultContext ctx;
        for (const auto& msg : msgs) {
            if (msg->returnvalue() == MIGRATED_FUNCTION_RETURN_VALUE) {
                continue;
            }
            std::vector<std::string> toNotify;
            try {
                recordResultLocked(msg, toNotify, &ctx);
            } catch (const std::exception& e) {
                SPDLOG_ERROR("Planner could not record the result of message {}: {}", msg->id(), e.what());
            }
            if (!toNotify.empty()) {
                notify.emplace_back(msg, std::move(toNotify));
            }
        }
    }
    for (const auto& [msg, hosts] : notify) {
        for (const auto& host : hosts) {
            faabric::scheduler::getFunctionCallClient(host)->setMessageResult(msg);
        }
    }
}

// Flat combining: whoever finds nobody draining takes the whole pending list
// through the planner in one go; everybody else just leaves their result
void Planner::submitMessageResult(std::shared_ptr<faabric::Message> msg)
{
    {
        std::lock_guard<std::mutex> lk(pendingResultsMx);
        pendingResults.push_back(std::move(msg));
        if (drainingResults) {
            return;
        }
        drainingResults = true;
    }
    std::vector<std::shared_ptr<faabric::Message>> batch;
    while (true) {
        batch.clear();
        {
            std::lock_guard<std::mutex> lk(pendingResultsMx);
            if (pendingResults.empty()) {
                drainingResults = false;
                return;
            }
            batch.swap(pendingResults);
        }
        setMessageResults(batch);
    }
}

void Planner::recordResultLocked(const std::shared_ptr<faabric::Message>& msg,
                                 std::vector<std::string>& toNotify,
                                 ResultContext* ctx)
{
    int appId = msg->appid();
    int msgId = msg->id();
    ResultContext local;
    if (ctx == nullptr) {
        ctx = &local;
    }
    if (!ctx->valid || ctx->appId != appId) {
        ctx->valid = true;
        ctx->appId = appId;
        ctx->results = &state.appResults[appId];
        ctx->hostName.clear();
        ctx->host = nullptr;
    }
    if (ctx->host == nullptr || ctx->hostName != msg->executedhost()) {
        auto found = state.hostMap.find(msg->executedhost());
        ctx->hostName = msg->executedhost();
        ctx->host = found == state.hostMap.end() ? nullptr : found->second;
        ctx->hostKnown = found != state.hostMap.end();
    }
    bool isFrozen = msg->returnvalue() == FROZEN_FUNCTION_RETURN_VALUE;
    if (isFrozen) {
        auto ev = state.evictedRequests.find(appId);
        if (ev == state.evictedRequests.end()) {
            SPDLOG_ERROR("Message {} is frozen but app {} not in map!", msgId, appId);
            throw std::runtime_error("Orphaned frozen message!");
        }
        // Remember where to resume from
        for (int i = 0; i < ev->second->messages_size(); i++) {
            auto* m = ev->second->mutable_messages(i);
            if (m->id() == msgId) {
                m->set_funcptr(msg->funcptr());
                m->set_inputdata(msg->inputdata());
                m->set_snapshotkey(msg->snapshotkey());
                m->set_returnvalue(msg->returnvalue());
                break;
            }
        }
    }
    const bool hostKnown = ctx->hostKnown && ctx->host != nullptr;
    auto& appResultsOfApp = *ctx->results;
    auto slot = appResultsOfApp.find(msgId);
    bool firstResult = slot == appResultsOfApp.end();
    if (hostKnown && (firstResult || isFrozen)) {
        releaseHostSlots(ctx->host);
    }
    if (!isFrozen) {
        if (firstResult) {
            appResultsOfApp.emplace_hint(appResultsOfApp.end(), msgId, msg);
        } else {
            slot->second = msg;
        }
    }
    auto inFlight = state.inFlightReqs.find(appId);
    if (inFlight != state.inFlightReqs.end()) {
        auto& req = inFlight->second.first;
        auto& decision = inFlight->second.second;
        auto& done = state.finishedInFlight[appId];
        // Ids are plain ints: a linear look-up is cheap, moving Message
        // objects around for every result is not
        // position of the message in the decision: hashed once per app (a
        // linear look-up per result is quadratic for a 1024-way fan-out)
        const auto& ids = decision->messageIds;
        auto& posOf = state.inFlightIdPos[appId];
        if (posOf.size() != ids.size()) {
            posOf.clear();
            posOf.reserve(ids.size());
            for (size_t k = 0; k < ids.size(); k++) {
                posOf.emplace(ids[k], (int)k);
            }
        }
        auto posIt = posOf.find(msgId);
        auto pos = posIt == posOf.end() ? ids.end() : ids.begin() + posIt->second;
        if (pos != ids.end() && *pos != msgId) {
            // the decision was edited (message removed / reordered): rebuild
            posOf.clear();
            pos = std::find(ids.begin(), ids.end(), msgId);
        }
        if (pos != ids.end() && done.insert(msgId).second) {
            int port = decision->mpiPorts.at((size_t)(pos - ids.begin()));
            if (hostKnown) {
                releaseHostMpiPort(ctx->host, port);
            }
            if ((int)done.size() == req->messages_size()) {
                SPDLOG_DEBUG("Planner removing app {} from in-flight", appId);
                state.inFlightReqs.erase(inFlight);
                state.finishedInFlight.erase(appId);
                state.inFlightIdPos.erase(appId);
                state.preloadedSchedulingDecisions.erase(appId);
                appFinishedCv.notify_all();
            }
        }
    }
    if (isFrozen) {
        return;
    }
    auto w = state.appResultWaiters.find(msgId);
    if (w != state.appResultWaiters.end()) {
        toNotify = std::move(w->second);
        state.appResultWaiters.erase(w);
    }
}

std::shared_ptr<faabric::Message> Planner::getMessageResult(
  std::shared_ptr<faabric::Message> msg)
{
    int appId = msg->appid();
    int msgId = msg->id();
    {
        std::shared_lock<std::shared_mutex> lock(plannerMx);
        auto a = state.appResults.find(appId);
        if (a != state.appResults.end()) {
            auto m = a->second.find(msgId);
            if (m != a->second.end()) {
                return m->second;
            }
        }
    }
    // Not there yet: remember who to call back (if they said who they are)
    if (!msg->mainhost().empty()) {
        std::unique_lock<std::shared_mutex> lock(plannerMx);
        auto a = state.appResults.find(appId);
        if (a != state.appResults.end()) {
            auto m = a->second.find(msgId);
            if (m != a->second.end()) {
                return m->second;
            }
        }
        state.appResultWaiters[msgId].push_back(msg->mainhost());
    }
    return nullptr;
}

void Planner::preloadSchedulingDecision(
  int32_t appId,
  std::shared_ptr<batch_scheduler::SchedulingDecision> decision)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    if (state.preloadedSchedulingDecisions.count(appId) > 0) {
        SPDLOG_ERROR("ERROR: preloaded scheduling decisions already contain app {}", appId);
        return;
    }
    SPDLOG_INFO("Pre-loading scheduling decision for app {}", appId);
    state.preloadedSchedulingDecisions[appId] = decision;
}

std::shared_ptr<batch_scheduler::SchedulingDecision>
Planner::getPreloadedSchedulingDecision(int32_t appId,
                                        std::shared_ptr<BatchExecuteRequest> ber)
{
    auto full = state.preloadedSchedulingDecisions.at(appId);
    // Only the group idxs present in this request
    auto filtered = std::make_shared<SchedulingDecision>(full->appId, full->groupId);
    for (const auto& msg : ber->messages()) {
        auto it = std::find(full->groupIdxs.begin(), full->groupIdxs.end(), msg.groupidx());
        if (it == full->groupIdxs.end()) {
            SPDLOG_ERROR("Group idx {} of app {} missing in preloaded decision", msg.groupidx(), appId);
            throw std::runtime_error("Group idx missing in preloaded decision");
        }
        size_t idx = (size_t)(it - full->groupIdxs.begin());
        filtered->addMessage(full->hosts.at(idx), msg.id(), full->appIdxs.at(idx), full->groupIdxs.at(idx));
        filtered->mpiPorts.back() = full->mpiPorts.at(idx);
    }
    return filtered;
}

std::shared_ptr<faabric::BatchExecuteRequestStatus> Planner::getBatchResults(
  int32_t appId)
{
    auto status = faabric::util::batchExecStatusFactory(appId);
    std::shared_ptr<BatchExecuteRequest> toThaw = nullptr;
    {
        std::shared_lock<std::shared_mutex> lock(plannerMx);
        auto ev = state.evictedRequests.find(appId);
        bool frozen = false;
        if (ev != state.evictedRequests.end()) {
            frozen = true;
            for (const auto& m : ev->second->messages()) {
                if (m.returnvalue() != FROZEN_FUNCTION_RETURN_VALUE) {
                    frozen = false; // still freezing
                }
            }
        }
        if (frozen) {
            if (state.inFlightReqs.count(appId) == 0) {
                toThaw = std::make_shared<BatchExecuteRequest>(*ev->second);
            }
            status->set_finished(false);
        } else {
            auto a = state.appResults.find(appId);
            if (a == state.appResults.end()) {
                return nullptr;
            }
            for (const auto& [id, m] : a->second) {
                *status->add_messageresults() = *m;
            }
            status->set_finished(state.inFlightReqs.count(appId) == 0);
        }
    }
    if (toThaw != nullptr) {
        // Polling the status of a frozen app doubles as the thaw trigger
        auto decision = callBatch(toThaw);
        if (*decision == NOT_ENOUGH_SLOTS_DECISION) {
            SPDLOG_DEBUG("Can not un-freeze app {} yet", appId);
        }
    }
    return status;
}

void Planner::compactInFlightLocked()
{
    for (auto& [appId, done] : state.finishedInFlight) {
        auto it = state.inFlightReqs.find(appId);
        if (it == state.inFlightReqs.end() || done.empty()) {
            continue;
        }
        auto& req = it->second.first;
        auto oldDecision = it->second.second;
        auto newDecision = std::make_shared<SchedulingDecision>(oldDecision->appId, oldDecision->groupId);
        newDecision->returnHost = oldDecision->returnHost;
        faabric::proto::RepeatedField<faabric::Message> kept;
        for (int i = 0; i < req->messages_size(); i++) {
            if (done.count(req->messages(i).id()) == 0) {
                *kept.Add() = std::move(*req->mutable_messages(i));
            }
        }
        for (int i = 0; i < oldDecision->nFunctions; i++) {
            if (done.count(oldDecision->messageIds[i]) == 0) {
                newDecision->addMessageInPosition(newDecision->nFunctions,
                                                  oldDecision->hosts[i],
                                                  oldDecision->messageIds[i],
                                                  oldDecision->appIdxs[i],
                                                  oldDecision->groupIdxs[i],
                                                  oldDecision->mpiPorts[i]);
            }
        }
        *req->mutable_messages() = std::move(kept);
        // Holders of the old decision object keep a consistent (stale) view
        it->second.second = newDecision;
    }
    state.finishedInFlight.clear();
    state.inFlightIdPos.clear();
}

bool Planner::waitForAppToFinish(int32_t appId, int timeoutMs)
{
    std::shared_lock<std::shared_mutex> lock(plannerMx);
    return appFinishedCv.wait_for(lock, std::chrono::milliseconds(timeoutMs), [&] {
        return state.inFlightReqs.find(appId) == state.inFlightReqs.end();
    });
}

std::shared_ptr<SchedulingDecision> Planner::getSchedulingDecision(
  std::shared_ptr<BatchExecuteRequest> req)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    compactInFlightLocked();
    auto it = state.inFlightReqs.find(req->appid());
    return it == state.inFlightReqs.end() ? nullptr : it->second.second;
}

faabric::batch_scheduler::InFlightReqs Planner::getInFlightReqs()
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    compactInFlightLocked();
    faabric::batch_scheduler::InFlightReqs copy;
    for (const auto& [appId, pair] : state.inFlightReqs) {
        copy[appId] = std::make_pair(std::make_shared<BatchExecuteRequest>(*pair.first),
                                     std::make_shared<SchedulingDecision>(*pair.second));
    }
    return copy;
}

int Planner::getNumMigrations()
{
    return state.numMigrations.load(std::memory_order_acquire);
}

std::set<std::string> Planner::getNextEvictedHostIps()
{
    std::shared_lock<std::shared_mutex> lock(plannerMx);
    return state.nextEvictedHostIps;
}

std::map<int32_t, std::shared_ptr<BatchExecuteRequest>> Planner::getEvictedReqs()
{
    std::shared_lock<std::shared_mutex> lock(plannerMx);
    std::map<int32_t, std::shared_ptr<BatchExecuteRequest>> out;
    for (const auto& [appId, ber] : state.evictedRequests) {
        out[appId] = std::make_shared<BatchExecuteRequest>(*ber);
    }
    return out;
}

std::string Planner::stateMain(const std::string& user, const std::string& key, const std::string& host, bool claim, bool drop)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    std::string lookup = user + "_" + key;
    if (drop) {
        state.stateMains.erase(lookup);
        return "";
    }
    auto it = state.stateMains.find(lookup);
    if (it != state.stateMains.end()) {
        return it->second;
    }
    if (!claim) {
        return "";
    }
    state.stateMains[lookup] = host;
    return host;
}

void Planner::setNextEvictedVm(const std::set<std::string>& vmIps)
{
    std::unique_lock<std::shared_mutex> lock(plannerMx);
    if (state.policy != "spot") {
        SPDLOG_ERROR("Error setting evicted VM with policy {} (must be spot)", state.policy);
        throw std::runtime_error("Error setting the next evicted VM!");
    }
    state.nextEvictedHostIps = vmIps;
}

// ---------------------------------------------------------------------------
// callBatch
// ---------------------------------------------------------------------------
std::shared_ptr<SchedulingDecision> Planner::callBatch(
  std::shared_ptr<BatchExecuteRequest> req)
{
    int appId = req->appid();
    std::shared_ptr<SchedulingDecision> decision;
    DecisionType type;
    {
        std::unique_lock<std::shared_mutex> lock(plannerMx);
        compactInFlightLocked();
        auto scheduler = faabric::batch_scheduler::getBatchScheduler();
        type = scheduler->getDecisionType(state.inFlightReqs, req);
        auto hostMapCopy = toSchedulerHostMap(state.hostMap, state.nextEvictedHostIps);
        const bool isNew = type == DecisionType::NEW;
        const bool isScale = type == DecisionType::SCALE_CHANGE;
        const bool isDist = type == DecisionType::DIST_CHANGE;
        const bool hasPreloaded = state.preloadedSchedulingDecisions.count(appId) > 0;

        // ---- elastic OpenMP scale-up: grow into idle slots of the main host
        if (isScale && req->elasticscalehint() && !hasPreloaded) {
            auto oldDec = state.inFlightReqs.at(appId).second;
            const std::string mainHost = oldDec->hosts.at(0);
            int avail = availableOpenMpSlots(appId, mainHost, state.hostMap, state.inFlightReqs);
            int requested = req->messages_size();
            int lastIdx = requested == 0 ? 0 : req->messages(requested - 1).groupidx();
            for (int k = 0; k < avail - requested; k++) {
                int idx = lastIdx + k + 1;
                faabric::Message* m = req->add_messages();
                if (requested == 0) {
                    *m = state.inFlightReqs.at(appId).first->messages(0);
                    m->set_mainhost(mainHost);
                    m->set_funcptr(req->groupid());
                } else {
                    *m = req->messages(requested - 1);
                }
                m->set_appidx(idx);
                m->set_groupidx(idx);
                m->set_id((int32_t)faabric::util::generateGid());
            }
            if (avail > requested) {
                SPDLOG_INFO("Elastically scaled-up app {} ({} -> {})", appId, requested, avail);
            }
        }

        // ---- a migration request is re-scheduled with the live messages
        if (isDist) {
            auto oldReq = state.inFlightReqs.at(appId).first;
            req->set_subtype(oldReq->subtype());
            req->clear_messages();
            for (const auto& m : oldReq->messages()) {
                *req->add_messages() = m;
            }
        }

        const bool isMpi = req->messages_size() > 0 && req->messages(0).ismpi();
        const bool isOmp = req->messages_size() > 0 && req->messages(0).isomp();

        // ---- OpenMP: slots announced by other apps but not yet occupied
        if (isOmp) {
            for (const auto& [otherId, pair] : state.inFlightReqs) {
                if (otherId == appId || pair.first->messages_size() == 0) {
                    continue;
                }
                int reserved = pair.first->messages(0).ompnumthreads() - pair.first->messages_size();
                if (reserved > 0 && !pair.second->hosts.empty()) {
                    auto h = hostMapCopy.find(pair.second->hosts.at(0));
                    if (h != hostMapCopy.end()) {
                        h->second->usedSlots += reserved;
                    }
                }
            }
        }

        // ---- make the decision
        std::shared_ptr<BatchExecuteRequest> knownSizeReq = nullptr;
        if (!isDist && hasPreloaded) {
            decision = getPreloadedSchedulingDecision(appId, req);
            if (isScale) {
                // The preloaded placement has now been consumed entirely
                state.preloadedSchedulingDecisions.erase(appId);
            }
        } else if (isNew && (isMpi || isOmp)) {
            // The app tells us its final size up-front: place ALL of it now,
            // dispatch only what we were given, preload the rest
            knownSizeReq = std::make_shared<BatchExecuteRequest>(*req);
            int finalSize = isMpi ? req->messages(0).mpiworldsize() : req->messages(0).ompnumthreads();
            for (int i = req->messages_size(); i < finalSize; i++) {
                faabric::Message* m = knownSizeReq->add_messages();
                m->set_appid(appId);
                m->set_groupidx(i);
            }
            decision = scheduler->makeSchedulingDecision(hostMapCopy, state.inFlightReqs, knownSizeReq);
        } else {
            decision = scheduler->makeSchedulingDecision(hostMapCopy, state.inFlightReqs, req);
        }

        // ---- sentinels
        if (*decision == NOT_ENOUGH_SLOTS_DECISION) {
            SPDLOG_ERROR("Not enough free slots to schedule app: {} (requested: {})", appId, req->messages_size());
            SPDLOG_DEBUG("{}", hostStateString(state.hostMap));
            return decision;
        }
        if (*decision == DO_NOT_MIGRATE_DECISION) {
            SPDLOG_INFO("Decided to not migrate app: {}", appId);
            return decision;
        }
        if (*decision == MUST_FREEZE_DECISION) {
            SPDLOG_INFO("Decided to FREEZE app: {}", appId);
            state.evictedRequests[appId] =
              std::make_shared<BatchExecuteRequest>(*state.inFlightReqs.at(appId).first);
            return decision;
        }

        if (!decision->isSingleHost() && req->singlehosthint()) {
            // A single-host app that does not fit on one host waits
            if (!(isNew && isOmp && req->elasticscalehint())) {
                SPDLOG_ERROR("User provided single-host hint in BER, but decision is not!");
            }
            return std::make_shared<SchedulingDecision>(NOT_ENOUGH_SLOTS_DECISION);
        }

        // ---- thawing a frozen app
        auto ev = state.evictedRequests.find(appId);
        if (ev != state.evictedRequests.end()) {
            if (isNew && isMpi) {
                SPDLOG_INFO("Decided to un-FREEZE app {}", appId);
                // Rank 0 restarts the world; the rest follow as a scale change
                faabric::Message first = req->messages(0);
                req->clear_messages();
                *req->add_messages() = first;
            } else if (isMpi && !isDist) {
                for (int i = 0; i < req->messages_size(); i++) {
                    for (int j = 1; j < ev->second->messages_size(); j++) {
                        const auto& frozen = ev->second->messages(j);
                        if (req->messages(i).groupidx() == frozen.groupidx()) {
                            auto* m = req->mutable_messages(i);
                            m->set_id(frozen.id());
                            m->set_funcptr(frozen.funcptr());
                            m->set_inputdata(frozen.inputdata());
                            m->set_snapshotkey(frozen.snapshotkey());
                            break;
                        }
                    }
                }
                state.evictedRequests.erase(ev);
            }
        }

        // ---- commit: new group id, claim slots/ports, update in-flight
        const bool skipClaim = decision->groupId == FIXED_SIZE_PRELOADED_DECISION_GROUPID;
        int newGroupId = (int)faabric::util::generateGid();
        decision->groupId = newGroupId;
        faabric::util::updateBatchExecGroupId(req, newGroupId);
        auto& broker = faabric::transport::getPointToPointBroker();

        switch (type) {
            case DecisionType::NEW: {
                for (size_t i = 0; i < decision->hosts.size(); i++) {
                    auto h = state.hostMap.at(decision->hosts.at(i));
                    claimHostSlots(h);
                    try {
                        decision->mpiPorts.at(i) = claimHostMpiPort(h);
                    } catch (const std::exception&) {
                        SPDLOG_ERROR("Error claiming MPI ports for app {}", appId);
                    }
                }
                if (knownSizeReq != nullptr) {
                    auto full = std::make_shared<SchedulingDecision>(*decision);
                    full->groupId = FIXED_SIZE_PRELOADED_DECISION_GROUPID;
                    state.preloadedSchedulingDecisions[appId] = full;
                    // Only the messages we were actually given run now
                    for (size_t i = (size_t)req->messages_size(); i < full->messageIds.size(); i++) {
                        decision->removeMessage(full->messageIds.at(i));
                    }
                }
                state.inFlightReqs[appId] = std::make_pair(req, decision);
                broker.setAndSendMappingsFromSchedulingDecision(*decision);
                break;
            }
            case DecisionType::SCALE_CHANGE: {
                auto oldReq = state.inFlightReqs.at(appId).first;
                auto oldDec = state.inFlightReqs.at(appId).second;
                faabric::util::updateBatchExecGroupId(oldReq, newGroupId);
                oldDec->groupId = newGroupId;
                for (int i = 0; i < req->messages_size(); i++) {
                    auto h = state.hostMap.at(decision->hosts.at(i));
                    *oldReq->add_messages() = req->messages(i);
                    oldDec->addMessage(decision->hosts.at(i), req->messages(i));
                    if (!skipClaim) {
                        claimHostSlots(h);
                        oldDec->mpiPorts.back() = claimHostMpiPort(h);
                    } else {
                        oldDec->mpiPorts.back() = decision->mpiPorts.at(i);
                    }
                }
                // Everybody (old and new members) gets the grown group
                broker.setAndSendMappingsFromSchedulingDecision(*oldDec);
                break;
            }
            case DecisionType::DIST_CHANGE: {
                auto oldReq = state.inFlightReqs.at(appId).first;
                auto oldDec = state.inFlightReqs.at(appId).second;
                std::set<std::string> oldHosts(oldDec->hosts.begin(), oldDec->hosts.end());
                std::set<std::string> newHosts(decision->hosts.begin(), decision->hosts.end());
                std::set<std::string> vacated;
                std::set_difference(oldHosts.begin(),
                                    oldHosts.end(),
                                    newHosts.begin(),
                                    newHosts.end(),
                                    std::inserter(vacated, vacated.begin()));
                SPDLOG_INFO("Decided to migrate app {}!", appId);
                oldDec->print("info");
                for (size_t i = 0; i < oldDec->hosts.size(); i++) {
                    if (decision->hosts.at(i) == oldDec->hosts.at(i)) {
                        continue;
                    }
                    auto from = state.hostMap.at(oldDec->hosts.at(i));
                    releaseHostSlots(from);
                    releaseHostMpiPort(from, oldDec->mpiPorts.at(i));
                    auto to = state.hostMap.at(decision->hosts.at(i));
                    claimHostSlots(to);
                    try {
                        decision->mpiPorts.at(i) = claimHostMpiPort(to);
                    } catch (const std::exception&) {
                        SPDLOG_ERROR("Error claiming MPI ports for app {}", appId);
                    }
                }
                decision->print("info");
                state.numMigrations += 1;
                faabric::util::updateBatchExecGroupId(oldReq, newGroupId);
                state.inFlightReqs.at(appId) = std::make_pair(oldReq, decision);
                broker.setAndSendMappingsFromSchedulingDecision(*decision);
                // Hosts the app leaves also need the new mappings
                broker.sendMappingsFromSchedulingDecision(*decision, vacated);
                break;
            }
            default:
                SPDLOG_ERROR("Unrecognised decision type: {} (app: {})", (int)type, appId);
                throw std::runtime_error("Unrecognised decision type");
        }
        // A migration is carried out by the app itself at its next migration
        // point; everything else is dispatched now.  Dispatch happens under
        // the lock so results cannot overtake the in-flight bookkeeping.
        // The in-flight table keeps mutating `decision` as results arrive, so
        // the caller gets its own snapshot taken before anything can finish
        auto returned = std::make_shared<batch_scheduler::SchedulingDecision>(*decision);
        if (type != DecisionType::DIST_CHANGE) {
            dispatchSchedulingDecision(req, decision);
        }
        return returned;
    }
}

void Planner::dispatchSchedulingDecision(
  std::shared_ptr<faabric::BatchExecuteRequest> req,
  std::shared_ptr<SchedulingDecision> decision)
{
    std::map<std::string, std::shared_ptr<BatchExecuteRequest>> perHost;
    const bool singleHost = decision->isSingleHost();
    for (int i = 0; i < req->messages_size(); i++) {
        const std::string& host = decision->hosts.at(i);
        auto& hr = perHost[host];
        if (hr == nullptr) {
            hr = std::make_shared<BatchExecuteRequest>();
            hr->set_appid((int32_t)decision->appId);
            hr->set_groupid(decision->groupId);
            hr->set_user(req->user());
            hr->set_function(req->function());
            hr->set_snapshotkey(req->snapshotkey());
            hr->set_type(req->type());
            hr->set_subtype(req->subtype());
            hr->set_contextdata(req->contextdata());
            hr->set_singlehost(singleHost);
            hr->set_singlehosthint(req->singlehosthint());
            hr->set_elasticscalehint(req->elasticscalehint());
        }
        *hr->add_messages() = req->messages(i);
    }
    const bool isThreads = req->type() == BatchExecuteRequest::THREADS;
    for (const auto& [host, hr] : perHost) {
        SPDLOG_DEBUG("Dispatching {} messages of app {} to {}", hr->messages_size(), req->appid(), host);
        // THREADS spanning hosts: remote hosts need the main thread snapshot
        if (isThreads && !singleHost) {
            std::string key = faabric::util::getMainThreadSnapshotKey(hr->messages(0));
            try {
                if (snapshotRegistry.deviceDescriptorExists(key)) {
                    // device-resident image: forward the descriptor only
                    if (host != req->messages(0).mainhost()) {
                        faabric::snapshot::getSnapshotClient(host)->pushDeviceSnapshot(
                          key, snapshotRegistry.getDeviceDescriptor(key));
                    }
                    goto dispatchFunctions;
                }
                auto snap = snapshotRegistry.getSnapshot(key);
                if (host != req->messages(0).mainhost()) {
                    faabric::snapshot::getSnapshotClient(host)->pushSnapshot(key, snap);
                }
            } catch (const std::runtime_error&) {
                SPDLOG_ERROR("Snapshot {} not registered in planner!", key);
            }
        }
        // Functions resuming from a snapshot (migration / thaw)
        if (!isThreads && !hr->messages(0).snapshotkey().empty()) {
            for (int i = 0; i < hr->messages_size(); i++) {
                const std::string& key = hr->messages(i).snapshotkey();
                try {
                    auto snap = snapshotRegistry.getSnapshot(key);
                    faabric::snapshot::getSnapshotClient(host)->pushSnapshot(key, snap);
                } catch (const std::runtime_error&) {
                    SPDLOG_ERROR("Snapshot {} not registered in planner!", key);
                }
            }
        }
    dispatchFunctions:
        faabric::scheduler::getFunctionCallClient(host)->executeFunctions(hr);
    }
}

Planner& getPlanner()
{
    static Planner planner;
    return planner;
}

} // namespace faabric::planner
