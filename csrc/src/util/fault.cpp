#include <faabric/util/fault.h>
#include <faabric/util/logging.h>
#include <faabric/util/string_tools.h>

#include <cstdlib>

namespace faabric::util {

FaultInjector::FaultInjector()
{
    const char* env = getenv("FAABRIC_FAULTS");
    if (env != nullptr && env[0] != '\0') {
        addRulesFromString(env);
    }
}

FaultInjector& FaultInjector::get()
{
    static FaultInjector f;
    return f;
}

void FaultInjector::addRule(const FaultRule& rule)
{
    std::lock_guard<std::mutex> lk(mx);
    rules.push_back(rule);
    nArmed.store((int)rules.size());
}

void FaultInjector::addRulesFromString(const std::string& spec)
{
    for (const auto& one : splitString(spec, ';')) {
        size_t colon = one.find(':');
        std::string action = toLower(trim(one.substr(0, colon)));
        FaultRule r;
        if (action == "drop") {
            r.action = FaultAction::DROP;
        } else if (action == "delay") {
            r.action = FaultAction::DELAY;
        } else if (action == "error") {
            r.action = FaultAction::ERROR;
        } else {
            SPDLOG_WARN("Ignoring unknown fault action '{}'", action);
            continue;
        }
        if (colon != std::string::npos) {
            for (const auto& kv : splitString(one.substr(colon + 1), ',')) {
                size_t eq = kv.find('=');
                if (eq == std::string::npos) {
                    continue;
                }
                std::string k = toLower(trim(kv.substr(0, eq)));
                int v = std::atoi(kv.substr(eq + 1).c_str());
                if (k == "port") {
                    r.port = v;
                } else if (k == "header") {
                    r.header = v;
                } else if (k == "ms") {
                    r.delayMs = v;
                } else if (k == "count") {
                    r.count = v;
                }
            }
        }
        addRule(r);
    }
}

void FaultInjector::clear()
{
    std::lock_guard<std::mutex> lk(mx);
    rules.clear();
    nArmed.store(0);
}

std::optional<FaultRule> FaultInjector::match(int port, int header)
{
    std::lock_guard<std::mutex> lk(mx);
    for (auto it = rules.begin(); it != rules.end(); ++it) {
        if ((it->port >= 0 && it->port != port) || (it->header >= 0 && it->header != header)) {
            continue;
        }
        FaultRule hit = *it;
        if (it->count > 0 && --it->count == 0) {
            rules.erase(it);
            nArmed.store((int)rules.size());
        }
        fired++;
        return hit;
    }
    return std::nullopt;
}

}
