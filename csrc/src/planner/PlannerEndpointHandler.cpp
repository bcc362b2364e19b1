#include <faabric/batch-scheduler/BatchScheduler.h>
#include <faabric/planner/Planner.h>
#include <faabric/planner/PlannerEndpointHandler.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/batch.h>
#include <faabric/util/json.h>
#include <faabric/util/logging.h>

namespace faabric::planner {

using faabric::endpoint::HttpRequest;
using faabric::endpoint::HttpResponse;
using faabric::util::JsonSerialisationException;

static void reply(HttpResponse& r, int status, std::string body)
{
    r.status = status;
    r.body = std::move(body);
}

// Parses `json` into `out`; on failure fills a 400 and returns false
template<typename M>
static bool parseOr400(const std::string& json, M* out, HttpResponse& r, const char* err)
{
    try {
        faabric::util::jsonToMessage(json, out);
        return true;
    } catch (JsonSerialisationException&) {
        reply(r, 400, err);
        return false;
    }
}

static void fillInFlight(GetInFlightAppsResponse& out)
{
    auto& planner = getPlanner();
    for (const auto& [appId, pair] : planner.getInFlightReqs()) {
        const auto& req = pair.first;
        const auto& decision = pair.second;
        auto* app = out.add_apps();
        app->set_appid(appId);
        app->set_subtype(req->subtype());
        const auto& first = req->messages(0);
        if (first.ismpi()) {
            app->set_size(first.mpiworldsize());
        }
        if (first.isomp()) {
            // An elastically scaled-up OpenMP app reports its actual size
            bool scaledUp = req->elasticscalehint() && first.ompnumthreads() < req->messages_size();
            app->set_size(scaledUp ? req->messages_size() : first.ompnumthreads());
        }
        for (const auto& ip : decision->hosts) {
            app->add_hostips(ip);
        }
    }
    out.set_nummigrations(planner.getNumMigrations());
    for (const auto& ip : planner.getNextEvictedHostIps()) {
        out.add_nextevictedvmips(ip);
    }
    for (const auto& [appId, ber] : planner.getEvictedReqs()) {
        auto* frozen = out.add_frozenapps();
        frozen->set_appid(appId);
        if (ber->messages(0).ismpi()) {
            frozen->set_size(ber->messages(0).mpiworldsize());
        }
    }
}

void PlannerEndpointHandler::onRequest(const HttpRequest& request, HttpResponse& response)
{
    response.headers["Server"] = "Planner endpoint";
    response.headers["Access-Control-Allow-Origin"] = "*";
    response.headers["Access-Control-Allow-Methods"] = "GET,POST,PUT,OPTIONS";
    response.headers["Access-Control-Allow-Headers"] = "User-Agent,Content-Type";
    response.headers["Content-Type"] = "text/plain";

    if (request.body.empty()) {
        SPDLOG_ERROR("Planner handler received empty request");
        return reply(response, 400, "Empty request");
    }

    HttpMessage msg;
    if (!parseOr400(request.body, &msg, response, "Bad JSON in request body")) {
        return;
    }

    auto& planner = getPlanner();
    switch (msg.type()) {
        case HttpMessage::RESET: {
            bool ok = planner.reset();
            return reply(response, ok ? 200 : 500, ok ? "Planner fully reset!" : "Failed to reset planner");
        }
        case HttpMessage::FLUSH_AVAILABLE_HOSTS: {
            bool ok = planner.flush(FlushType::Hosts);
            return reply(response, ok ? 200 : 500, ok ? "Flushed available hosts!" : "Failed flushing available hosts!");
        }
        case HttpMessage::FLUSH_EXECUTORS: {
            bool ok = planner.flush(FlushType::Executors);
            return reply(response, ok ? 200 : 500, ok ? "Flushed executors!" : "Failed flushing executors!");
        }
        case HttpMessage::FLUSH_SCHEDULING_STATE: {
            planner.flush(FlushType::SchedulingState);
            return reply(response, 200, "Flushed scheduling state!");
        }
        case HttpMessage::GET_AVAILABLE_HOSTS: {
            AvailableHostsResponse hosts;
            for (auto& h : planner.getAvailableHosts()) {
                *hosts.add_hosts() = *h;
            }
            return reply(response, 200, faabric::util::messageToJson(hosts));
        }
        case HttpMessage::GET_CONFIG: {
            return reply(response, 200, faabric::util::messageToJson(planner.getConfig()));
        }
        case HttpMessage::GET_EXEC_GRAPH: {
            faabric::Message payload;
            if (!parseOr400(msg.payloadjson(), &payload, response, "Bad JSON in request body")) {
                return;
            }
            auto graph = faabric::util::getFunctionExecGraph(payload);
            if (graph.rootNode.msg.id() == 0) {
                SPDLOG_ERROR("Error processing GET_EXEC_GRAPH request");
                return reply(response, 500, "Failed getting exec. graph!");
            }
            return reply(response, 200, faabric::util::execGraphToJson(graph));
        }
        case HttpMessage::GET_IN_FLIGHT_APPS: {
            GetInFlightAppsResponse apps;
            fillInFlight(apps);
            return reply(response, 200, faabric::util::messageToJson(apps));
        }
        case HttpMessage::EXECUTE_BATCH: {
            auto ber = std::make_shared<faabric::BatchExecuteRequest>();
            if (!parseOr400(msg.payloadjson(), ber.get(), response, "Bad JSON in body's payload")) {
                return;
            }
            if (!faabric::util::isBatchExecRequestValid(ber)) {
                return reply(response, 400, "Bad BatchExecRequest");
            }
            auto decision = planner.callBatch(ber);
            if (*decision == NOT_ENOUGH_SLOTS_DECISION) {
                return reply(response, 500, "No available hosts");
            }
            auto status = faabric::util::batchExecStatusFactory(ber);
            return reply(response, 200, faabric::util::messageToJson(*status));
        }
        case HttpMessage::EXECUTE_BATCH_STATUS: {
            faabric::BatchExecuteRequestStatus asked;
            if (!parseOr400(msg.payloadjson(), &asked, response, "Bad JSON in request body")) {
                return;
            }
            auto actual = planner.getBatchResults(asked.appid());
            if (actual == nullptr) {
                return reply(response, 500, "App not registered in results");
            }
            return reply(response, 200, faabric::util::messageToJson(*actual));
        }
        case HttpMessage::PRELOAD_SCHEDULING_DECISION: {
            faabric::BatchExecuteRequest ber;
            if (!parseOr400(msg.payloadjson(), &ber, response, "Bad JSON in request body")) {
                return;
            }
            // The "BER" here is a carrier for (host, id, appIdx, groupIdx)
            auto decision =
              std::make_shared<faabric::batch_scheduler::SchedulingDecision>(ber.appid(), ber.groupid());
            for (const auto& m : ber.messages()) {
                decision->addMessage(m.executedhost(), m.id(), m.appidx(), m.groupidx());
            }
            planner.preloadSchedulingDecision(decision->appId, decision);
            return reply(response, 200, "Decision pre-loaded to planner");
        }
        case HttpMessage::SET_POLICY: {
            const std::string& policy = msg.payloadjson();
            try {
                planner.setPolicy(policy);
            } catch (std::exception&) {
                return reply(response, 400, "Unrecognised policy name: " + policy);
            }
            return reply(response, 200, "Policy set correctly");
        }
        case HttpMessage::GET_POLICY: {
            return reply(response, 200, planner.getPolicy());
        }
        case HttpMessage::SET_NEXT_EVICTED_VM: {
            SetEvictedVmIpsRequest evicted;
            if (!parseOr400(msg.payloadjson(), &evicted, response, "Bad JSON in body's payload")) {
                return;
            }
            std::set<std::string> ips(evicted.vmips().begin(), evicted.vmips().end());
            try {
                planner.setNextEvictedVm(ips);
            } catch (std::exception&) {
                return reply(response, 400, "Next evicted VM must only be set in 'spot' policy");
            }
            return reply(response, 200, "Next evicted VM set");
        }
        default: {
            SPDLOG_ERROR("Unrecognised message type {}", (int)msg.type());
            return reply(response, 400, "Unrecognised message type");
        }
    }
}

}
