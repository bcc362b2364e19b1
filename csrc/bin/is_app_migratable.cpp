// Migration oracle: would the given policy migrate app APP_ID under this
// cluster occupation?  Exit code 0 = yes, 1 = no.  Uses the very same
// BatchScheduler code as the planner (reference: src/planner/is_app_migratable.cpp).
//
//   is_app_migratable <bin-pack|compact|spot> <appId> <occupation.csv>
//
// CSV: a header line, then "workerIp,slot,slot,..." where each slot is the id
// of the app running there or -1 when free.
#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/util/batch.h>
#include <faabric/util/gids.h>
#include <faabric/util/logging.h>
#include <faabric/util/string_tools.h>

#include <fstream>
#include <iostream>

using namespace faabric::batch_scheduler;

static void loadOccupation(const std::string& path, HostMap& hosts, InFlightReqs& apps)
{
    std::ifstream in(path);
    if (!in.is_open()) {
        throw std::runtime_error("Cannot open occupation file " + path);
    }
    std::string line;
    std::getline(in, line); // header
    while (std::getline(in, line)) {
        auto cells = faabric::util::splitString(faabric::util::trim(line), ',');
        if (cells.empty()) {
            continue;
        }
        const std::string& ip = cells[0];
        auto host = std::make_shared<HostState>(ip, 0, 0);
        hosts[ip] = host;
        for (size_t i = 1; i < cells.size(); i++) {
            int appId = std::stoi(cells[i]);
            host->slots++;
            if (appId == -1) {
                continue;
            }
            host->usedSlots++;
            auto it = apps.find(appId);
            if (it == apps.end()) {
                auto ber = faabric::util::batchExecFactory();
                ber->set_user("foo");
                ber->set_function("bar");
                faabric::util::updateBatchExecAppId(ber, appId);
                faabric::util::updateBatchExecGroupId(ber, (int)faabric::util::generateGid());
                auto decision = std::make_shared<SchedulingDecision>(ber->appid(), ber->groupid());
                it = apps.emplace(appId, std::make_pair(ber, decision)).first;
            }
            auto& [ber, decision] = it->second;
            auto* msg = ber->add_messages();
            msg->set_user(ber->user());
            msg->set_function(ber->function());
            msg->set_appid(ber->appid());
            msg->set_groupid(ber->groupid());
            msg->set_id((int)faabric::util::generateGid());
            msg->set_groupidx((int)decision->hosts.size());
            msg->set_appidx((int)decision->hosts.size());
            decision->addMessage(ip, *msg);
        }
    }
}

int main(int argc, char** argv)
{
    if (argc != 4) {
        std::cerr << "usage: is_app_migratable <policy> <appId> <occupation.csv>" << std::endl;
        return 2;
    }
    std::string policy = argv[1];
    int appId = std::atoi(argv[2]);
    faabric::util::initLogging();

    HostMap hosts;
    InFlightReqs apps;
    loadOccupation(argv[3], hosts, apps);
    if (apps.find(appId) == apps.end()) {
        std::cerr << "App " << appId << " is not running anywhere" << std::endl;
        return 2;
    }
    for (const auto& [ip, host] : hosts) {
        std::cout << "IP: " << ip << " - Slots: " << host->usedSlots << "/" << host->slots << std::endl;
    }
    for (const auto& [id, pair] : apps) {
        std::cout << "App " << id << " runs " << pair.second->hosts.size() << " messages" << std::endl;
    }

    resetBatchScheduler(policy);
    auto req = apps.at(appId).first;
    req->set_type(faabric::BatchExecuteRequest::MIGRATION);
    auto decision = getBatchScheduler()->makeSchedulingDecision(hosts, apps, req);
    if (*decision == DO_NOT_MIGRATE_DECISION) {
        std::cout << "NOT migrating app: " << appId << std::endl;
        return 1;
    }
    std::cout << "Migrating app: " << appId << std::endl;
    decision->print("info");
    return 0;
}
