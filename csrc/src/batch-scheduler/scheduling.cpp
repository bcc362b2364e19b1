#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/batch-scheduler/DecisionCache.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/ptp.h>

#include <mutex>
#include <sstream>

namespace faabric::batch_scheduler {

// ---------------------------------------------------------------------------
// SchedulingDecision
// ---------------------------------------------------------------------------
SchedulingDecision::SchedulingDecision(uint32_t appIdIn, int32_t groupIdIn)
  : appId(appIdIn)
  , groupId(groupIdIn)
{}

bool SchedulingDecision::isSingleHost() const
{
    // All messages on the same host (an empty decision trivially so)
    return std::set<std::string>(hosts.begin(), hosts.end()).size() <= 1;
}

void SchedulingDecision::addMessage(const std::string& host,
                                    const faabric::Message& msg)
{
    addMessage(host, msg.id(), msg.appidx(), msg.groupidx());
}

void SchedulingDecision::addMessage(const std::string& host,
                                    int32_t messageId,
                                    int32_t appIdx)
{
    addMessage(host, messageId, appIdx, 0);
}

void SchedulingDecision::addMessage(const std::string& host,
                                    int32_t messageId,
                                    int32_t appIdx,
                                    int32_t groupIdx)
{
    nFunctions++;
    hosts.emplace_back(host);
    messageIds.emplace_back(messageId);
    appIdxs.emplace_back(appIdx);
    groupIdxs.emplace_back(groupIdx);
    mpiPorts.emplace_back(0);
}

void SchedulingDecision::addMessageInPosition(int32_t pos,
                                              const std::string& host,
                                              int32_t messageId,
                                              int32_t appIdx,
                                              int32_t groupIdx,
                                              int32_t mpiPort)
{
    nFunctions++;
    int desired = std::max<int>(pos + 1, nFunctions);
    if ((int)hosts.size() < desired) {
        hosts.resize(desired);
        messageIds.resize(desired, 0);
        appIdxs.resize(desired, 0);
        groupIdxs.resize(desired, 0);
        mpiPorts.resize(desired, 0);
    }
    hosts.at(pos) = host;
    messageIds.at(pos) = messageId;
    appIdxs.at(pos) = appIdx;
    groupIdxs.at(pos) = groupIdx;
    mpiPorts.at(pos) = mpiPort;
}

int32_t SchedulingDecision::removeMessage(int32_t messageId)
{
    auto it = std::find(messageIds.begin(), messageIds.end(), messageId);
    if (it == messageIds.end()) {
        SPDLOG_ERROR("Message {} not part of decision for app {}", messageId, appId);
        throw std::runtime_error("Attempting to remove a message not in decision");
    }
    size_t idx = (size_t)(it - messageIds.begin());
    int32_t port = mpiPorts.at(idx);
    nFunctions--;
    hosts.erase(hosts.begin() + idx);
    messageIds.erase(messageIds.begin() + idx);
    appIdxs.erase(appIdxs.begin() + idx);
    groupIdxs.erase(groupIdxs.begin() + idx);
    mpiPorts.erase(mpiPorts.begin() + idx);
    return port;
}

std::set<std::string> SchedulingDecision::uniqueHosts()
{
    return std::set<std::string>(hosts.begin(), hosts.end());
}

SchedulingDecision SchedulingDecision::fromPointToPointMappings(
  faabric::PointToPointMappings& mappings)
{
    SchedulingDecision decision(mappings.appid(), mappings.groupid());
    for (const auto& m : mappings.mappings()) {
        decision.addMessage(m.host(), m.messageid(), m.appidx(), m.groupidx());
        decision.mpiPorts.back() = m.mpiport();
    }
    return decision;
}

std::string SchedulingDecision::toString() const
{
    std::ostringstream os;
    os << "-------------- Decision for App: " << appId << " ----------------\n";
    os << "MsgId\tAppId\tGroupId\tGrIdx\tHostIp\tPort\n";
    for (int i = 0; i < (int)hosts.size(); i++) {
        os << messageIds.at(i) << "\t" << appId << "\t" << groupId << "\t"
           << groupIdxs.at(i) << "\t" << hosts.at(i) << "\t" << mpiPorts.at(i)
           << "\n";
    }
    os << "------------- End Decision for App " << appId << " ---------------";
    return os.str();
}

void SchedulingDecision::print(const std::string& logLevel)
{
    std::string s = toString();
    if (logLevel == "info") {
        SPDLOG_INFO("{}", s);
    } else if (logLevel == "warn") {
        SPDLOG_WARN("{}", s);
    } else {
        SPDLOG_DEBUG("{}", s);
    }
}

} // namespace faabric::batch_scheduler

namespace faabric::util {
faabric::PointToPointMappings ptpMappingsFromSchedulingDecision(
  std::shared_ptr<faabric::batch_scheduler::SchedulingDecision> decision)
{
    faabric::PointToPointMappings mappings;
    mappings.set_appid((int32_t)decision->appId);
    mappings.set_groupid(decision->groupId);
    for (int i = 0; i < (int)decision->hosts.size(); i++) {
        auto* m = mappings.add_mappings();
        m->set_host(decision->hosts.at(i));
        m->set_messageid(decision->messageIds.at(i));
        m->set_appidx(decision->appIdxs.at(i));
        m->set_groupidx(decision->groupIdxs.at(i));
        m->set_mpiport(decision->mpiPorts.at(i));
    }
    return mappings;
}
}

namespace faabric::batch_scheduler {

// ---------------------------------------------------------------------------
// Scheduler selection
// ---------------------------------------------------------------------------
static std::shared_ptr<BatchScheduler> activeScheduler;
static std::mutex schedulerMx;

std::shared_ptr<BatchScheduler> getBatchScheduler()
{
    std::lock_guard<std::mutex> lk(schedulerMx);
    if (activeScheduler != nullptr) {
        return activeScheduler;
    }
    const std::string& mode = faabric::util::getSystemConfig().batchSchedulerMode;
    if (mode == "bin-pack") {
        activeScheduler = std::make_shared<BinPackScheduler>();
    } else if (mode == "compact") {
        activeScheduler = std::make_shared<CompactScheduler>();
    } else if (mode == "spot") {
        activeScheduler = std::make_shared<SpotScheduler>();
    } else {
        SPDLOG_ERROR("Unrecognised batch scheduler mode: {}", mode);
        throw std::runtime_error("Unrecognised batch scheduler mode");
    }
    return activeScheduler;
}

void resetBatchScheduler()
{
    std::lock_guard<std::mutex> lk(schedulerMx);
    activeScheduler = nullptr;
}

void resetBatchScheduler(const std::string& newMode)
{
    resetBatchScheduler();
    faabric::util::getSystemConfig().batchSchedulerMode = newMode;
    getBatchScheduler();
}

DecisionType BatchScheduler::getDecisionType(
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    if (inFlightReqs.find(req->appid()) == inFlightReqs.end()) {
        return DecisionType::NEW;
    }
    if (req->type() == faabric::BatchExecuteRequest::MIGRATION) {
        return DecisionType::DIST_CHANGE;
    }
    return DecisionType::SCALE_CHANGE;
}

// ---------------------------------------------------------------------------
// Greedy engine
// ---------------------------------------------------------------------------
std::map<std::string, int> GreedyPackScheduler::hostHistogram(
  const std::shared_ptr<SchedulingDecision>& decision)
{
    std::map<std::string, int> h;
    for (const auto& host : decision->hosts) {
        h[host]++;
    }
    return h;
}

std::set<std::string> GreedyPackScheduler::filterHosts(
  HostMap& hostMap,
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    return {};
}

std::vector<Host> GreedyPackScheduler::getSortedHosts(
  HostMap& hostMap,
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req,
  const DecisionType& decisionType)
{
    std::vector<Host> hosts;
    hosts.reserve(hostMap.size());
    for (auto& [ip, host] : hostMap) {
        hosts.push_back(host);
    }
    std::map<std::string, int> appFreq;
    if (decisionType != DecisionType::NEW) {
        appFreq = hostHistogram(inFlightReqs.at(req->appid()).second);
    }
    auto freqOf = [&appFreq](const Host& h) {
        auto it = appFreq.find(h->ip);
        return it == appFreq.end() ? 0 : it->second;
    };
    // Most free slots, then biggest, then highest address
    auto roomier = [](const Host& a, const Host& b) {
        int fa = numSlotsAvailable(a);
        int fb = numSlotsAvailable(b);
        if (fa != fb) {
            return fa > fb;
        }
        if (a->slots != b->slots) {
            return a->slots > b->slots;
        }
        return a->ip > b->ip;
    };
    auto byAppFreq = [&](const Host& a, const Host& b) {
        int qa = freqOf(a);
        int qb = freqOf(b);
        if (qa != qb) {
            return qa > qb;
        }
        return roomier(a, b);
    };

    switch (decisionType) {
        case DecisionType::NEW:
            std::sort(hosts.begin(), hosts.end(), roomier);
            break;
        case DecisionType::SCALE_CHANGE:
            // Prefer hosts that already run this app
            std::sort(hosts.begin(), hosts.end(), byAppFreq);
            break;
        case DecisionType::DIST_CHANGE: {
            // Pretend the app is not running: hand its slots back first
            for (auto& h : hosts) {
                int q = freqOf(h);
                if (q > 0) {
                    freeSlots(h, q);
                }
            }
            switch (migrationOrder()) {
                case MigrationOrder::MostFreeThenAppFrequency:
                    std::sort(hosts.begin(), hosts.end(), [&](const Host& a, const Host& b) {
                        int fa = numSlotsAvailable(a);
                        int fb = numSlotsAvailable(b);
                        if (fa != fb) {
                            return fa > fb;
                        }
                        return byAppFreq(a, b);
                    });
                    break;
                case MigrationOrder::FullestFirst:
                    std::sort(hosts.begin(), hosts.end(), [&](const Host& a, const Host& b) {
                        if (a->usedSlots != b->usedSlots) {
                            return a->usedSlots > b->usedSlots;
                        }
                        return roomier(a, b);
                    });
                    break;
                case MigrationOrder::AppFrequencyFirst:
                    std::sort(hosts.begin(), hosts.end(), byAppFreq);
                    break;
            }
            break;
        }
        default:
            SPDLOG_ERROR("Unrecognised decision type: {}", (int)decisionType);
            throw std::runtime_error("Unrecognised decision type");
    }
    return hosts;
}

bool GreedyPackScheduler::isFirstDecisionBetter(
  std::shared_ptr<SchedulingDecision> decisionA,
  std::shared_ptr<SchedulingDecision> decisionB)
{
    // (number of hosts, number of cross-host links in a fully connected app)
    auto score = [](const std::shared_ptr<SchedulingDecision>& d) {
        auto hist = hostHistogram(d);
        if (hist.size() <= 1) {
            return std::make_pair((int)hist.size(), 0);
        }
        long total = 0;
        for (const auto& [h, n] : hist) {
            total += n;
        }
        long links = 0;
        for (const auto& [h, n] : hist) {
            links += (long)n * (total - n);
        }
        return std::make_pair((int)hist.size(), (int)(links / 2));
    };
    auto a = score(decisionA);
    auto b = score(decisionB);
    if (a.first != b.first) {
        return a.first < b.first;
    }
    return a.second < b.second;
}

std::shared_ptr<SchedulingDecision> GreedyPackScheduler::minimiseNumOfMigrations(
  std::shared_ptr<SchedulingDecision> fresh,
  std::shared_ptr<SchedulingDecision> old)
{
    auto out = std::make_shared<SchedulingDecision>(old->appId, old->groupId);
    auto budget = hostHistogram(fresh);
    const int n = (int)old->hosts.size();
    std::vector<bool> placed(n, false);

    // Pass 1: whoever can stay where it is, stays (and keeps its MPI port)
    for (int i = 0; i < n; i++) {
        auto it = budget.find(old->hosts.at(i));
        if (it != budget.end() && it->second > 0) {
            out->addMessageInPosition(i,
                                      old->hosts.at(i),
                                      old->messageIds.at(i),
                                      old->appIdxs.at(i),
                                      old->groupIdxs.at(i),
                                      old->mpiPorts.at(i));
            it->second--;
            placed[i] = true;
        }
    }
    // Pass 2: the rest fill whatever capacity is left; their port is unknown
    // until the planner assigns one on the destination
    for (int i = 0; i < n; i++) {
        if (placed[i]) {
            continue;
        }
        auto it = std::find_if(budget.begin(), budget.end(), [](const auto& kv) {
            return kv.second > 0;
        });
        if (it == budget.end()) {
            throw std::runtime_error("No next host with slots found!");
        }
        out->addMessageInPosition(i,
                                  it->first,
                                  old->messageIds.at(i),
                                  old->appIdxs.at(i),
                                  old->groupIdxs.at(i),
                                  -1);
        it->second--;
    }
    return out;
}

std::shared_ptr<SchedulingDecision> GreedyPackScheduler::makeSchedulingDecision(
  HostMap& hostMap,
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    auto fresh = std::make_shared<SchedulingDecision>(req->appid(), 0);
    std::set<std::string> removed = filterHosts(hostMap, inFlightReqs, req);
    DecisionType type = getDecisionType(inFlightReqs, req);
    std::vector<Host> hosts = getSortedHosts(hostMap, inFlightReqs, req, type);

    if (honourSingleHostHint() && req->singlehosthint() &&
        req->messages_size() > 0 && req->messages(0).isomp() &&
        hosts.size() > 1) {
        hosts.resize(1);
    }

    int left = req->messages_size();
    int next = 0;
    for (auto& h : hosts) {
        int take = std::min<int>(left, numSlotsAvailable(h));
        for (int i = 0; i < take; i++) {
            fresh->addMessage(h->ip, req->messages(next++));
        }
        left -= take;
        if (left == 0) {
            break;
        }
    }

    if (type == DecisionType::DIST_CHANGE) {
        auto old = inFlightReqs.at(req->appid()).second;
        return resolveDistChange(hostMap, removed, fresh, old, left);
    }
    if (left > 0) {
        return std::make_shared<SchedulingDecision>(NOT_ENOUGH_SLOTS_DECISION);
    }
    return fresh;
}

// ---------------------------------------------------------------------------
// Bin-pack
// ---------------------------------------------------------------------------
std::shared_ptr<SchedulingDecision> BinPackScheduler::resolveDistChange(
  HostMap& hostMap,
  const std::set<std::string>& removedHosts,
  std::shared_ptr<SchedulingDecision> fresh,
  std::shared_ptr<SchedulingDecision> old,
  int numLeftToSchedule)
{
    if (numLeftToSchedule > 0) {
        return std::make_shared<SchedulingDecision>(NOT_ENOUGH_SLOTS_DECISION);
    }
    if (isFirstDecisionBetter(fresh, old)) {
        return minimiseNumOfMigrations(fresh, old);
    }
    return std::make_shared<SchedulingDecision>(DO_NOT_MIGRATE_DECISION);
}

// ---------------------------------------------------------------------------
// Compact
// ---------------------------------------------------------------------------
std::set<std::string> CompactScheduler::filterHosts(
  HostMap& hostMap,
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    // The tenant id travels in the BER's subType
    std::set<std::string> removed;
    int tenant = req->subtype();
    for (const auto& [appId, pair] : inFlightReqs) {
        if (pair.first->subtype() == tenant) {
            continue;
        }
        for (const auto& h : pair.second->hosts) {
            if (hostMap.erase(h) > 0) {
                removed.insert(h);
            }
        }
    }
    return removed;
}

bool CompactScheduler::isFirstDecisionBetter(
  std::shared_ptr<SchedulingDecision> decisionA,
  std::shared_ptr<SchedulingDecision> decisionB)
{
    throw std::runtime_error("Method not supported for COMPACT scheduler");
}

bool CompactScheduler::isFirstDecisionBetter(
  HostMap& hostMap,
  std::shared_ptr<SchedulingDecision> newDecision,
  std::shared_ptr<SchedulingDecision> oldDecision)
{
    // hostMap has the app's own slots already released: count how many hosts
    // would be completely idle under either placement
    auto idleHostsWith = [&hostMap](const std::shared_ptr<SchedulingDecision>& d) {
        std::map<std::string, int> used;
        for (const auto& [ip, h] : hostMap) {
            used[ip] = h->usedSlots;
        }
        for (const auto& ip : d->hosts) {
            auto it = used.find(ip);
            if (it == used.end()) {
                SPDLOG_ERROR("Host {} of decision missing from host map", ip);
                continue;
            }
            it->second++;
        }
        int idle = 0;
        for (const auto& [ip, n] : used) {
            idle += (n == 0) ? 1 : 0;
        }
        return idle;
    };
    return idleHostsWith(newDecision) > idleHostsWith(oldDecision);
}

std::shared_ptr<SchedulingDecision> CompactScheduler::resolveDistChange(
  HostMap& hostMap,
  const std::set<std::string>& removedHosts,
  std::shared_ptr<SchedulingDecision> fresh,
  std::shared_ptr<SchedulingDecision> old,
  int numLeftToSchedule)
{
    if (numLeftToSchedule > 0) {
        return std::make_shared<SchedulingDecision>(NOT_ENOUGH_SLOTS_DECISION);
    }
    if (isFirstDecisionBetter(hostMap, fresh, old)) {
        return minimiseNumOfMigrations(fresh, old);
    }
    return std::make_shared<SchedulingDecision>(DO_NOT_MIGRATE_DECISION);
}

// ---------------------------------------------------------------------------
// Spot
// ---------------------------------------------------------------------------
std::set<std::string> SpotScheduler::filterHosts(
  HostMap& hostMap,
  const InFlightReqs& inFlightReqs,
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    // Hosts about to be reclaimed are tainted by overwriting their address
    std::set<std::string> removed;
    for (auto it = hostMap.begin(); it != hostMap.end();) {
        if (it->second->ip == MUST_EVICT_IP) {
            removed.insert(it->first);
            it = hostMap.erase(it);
        } else {
            ++it;
        }
    }
    return removed;
}

bool SpotScheduler::isFirstDecisionBetter(
  std::shared_ptr<SchedulingDecision> decisionA,
  std::shared_ptr<SchedulingDecision> decisionB)
{
    throw std::runtime_error("Method not supported for SPOT scheduler");
}

std::shared_ptr<SchedulingDecision> SpotScheduler::resolveDistChange(
  HostMap& hostMap,
  const std::set<std::string>& removedHosts,
  std::shared_ptr<SchedulingDecision> fresh,
  std::shared_ptr<SchedulingDecision> old,
  int numLeftToSchedule)
{
    if (numLeftToSchedule > 0) {
        // Nowhere to run once the evicted hosts are gone: park the app
        return std::make_shared<SchedulingDecision>(MUST_FREEZE_DECISION);
    }
    bool touchesEvicted = std::any_of(
      old->hosts.begin(), old->hosts.end(), [&](const std::string& h) {
          return removedHosts.count(h) > 0;
      });
    if (touchesEvicted) {
        return minimiseNumOfMigrations(fresh, old);
    }
    return std::make_shared<SchedulingDecision>(DO_NOT_MIGRATE_DECISION);
}

// ---------------------------------------------------------------------------
// Decision cache
// ---------------------------------------------------------------------------
CachedDecision::CachedDecision(const std::vector<std::string>& hostsIn,
                               int groupIdIn)
  : hosts(hostsIn)
  , groupId(groupIdIn)
{}

std::string DecisionCache::getCacheKey(
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    return std::to_string(req->appid()) + "_" +
           std::to_string(req->messages_size());
}

std::shared_ptr<CachedDecision> DecisionCache::getCachedDecision(
  std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    std::string key = getCacheKey(req);
    std::shared_lock<std::shared_mutex> lk(mx);
    auto it = cachedDecisions.find(key);
    if (it == cachedDecisions.end()) {
        return nullptr;
    }
    // Sanity check: the cached placement must cover every message
    if ((int)it->second->getHosts().size() != req->messages_size()) {
        SPDLOG_ERROR("Cached decision for {} has wrong size", key);
        throw std::runtime_error("Invalid cached scheduling decision");
    }
    return it->second;
}

void DecisionCache::addCachedDecision(
  std::shared_ptr<faabric::BatchExecuteRequest> req,
  SchedulingDecision& decision)
{
    if ((int)decision.hosts.size() != req->messages_size()) {
        SPDLOG_ERROR("Trying to cache a decision of size {} for a request of size {}",
                     decision.hosts.size(),
                     req->messages_size());
        throw std::runtime_error("Invalid decision caching");
    }
    std::string key = getCacheKey(req);
    std::unique_lock<std::shared_mutex> lk(mx);
    if (cachedDecisions.count(key) > 0) {
        return;
    }
    cachedDecisions[key] =
      std::make_shared<CachedDecision>(decision.hosts, decision.groupId);
}

void DecisionCache::clear()
{
    std::unique_lock<std::shared_mutex> lk(mx);
    cachedDecisions.clear();
}

DecisionCache& getSchedulingDecisionCache()
{
    static DecisionCache c;
    return c;
}

} // namespace faabric::batch_scheduler
