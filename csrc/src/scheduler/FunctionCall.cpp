#include <faabric/planner/PlannerClient.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/scheduler/FunctionCallServer.h>
#include <faabric/scheduler/Scheduler.h>
#include <faabric/state/State.h>
#include <faabric/transport/common.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/testing.h>

namespace faabric::scheduler {

// ---------------------------------------------------------------------------
// Mock capture
// ---------------------------------------------------------------------------
static std::mutex mockMutex;
static std::vector<std::pair<std::string, faabric::Message>> functionCalls;
static std::vector<std::pair<std::string, faabric::EmptyRequest>> flushCalls;
static std::vector<std::pair<std::string, std::shared_ptr<faabric::BatchExecuteRequest>>> batchMessages;
static std::vector<std::pair<std::string, std::shared_ptr<faabric::Message>>> messageResults;

std::vector<std::pair<std::string, faabric::Message>> getFunctionCalls()
{
    std::lock_guard<std::mutex> lk(mockMutex);
    return functionCalls;
}

std::vector<std::pair<std::string, faabric::EmptyRequest>> getFlushCalls()
{
    std::lock_guard<std::mutex> lk(mockMutex);
    return flushCalls;
}

std::vector<std::pair<std::string, std::shared_ptr<faabric::BatchExecuteRequest>>> getBatchRequests()
{
    std::lock_guard<std::mutex> lk(mockMutex);
    return batchMessages;
}

std::vector<std::pair<std::string, std::shared_ptr<faabric::Message>>> getMessageResults()
{
    std::lock_guard<std::mutex> lk(mockMutex);
    return messageResults;
}

void clearMockRequests()
{
    std::lock_guard<std::mutex> lk(mockMutex);
    functionCalls.clear();
    flushCalls.clear();
    batchMessages.clear();
    messageResults.clear();
}

// ---------------------------------------------------------------------------
// Client pool: one client per (thread, host)
// ---------------------------------------------------------------------------
static thread_local std::unordered_map<std::string, std::shared_ptr<FunctionCallClient>> tlsClients;

std::shared_ptr<FunctionCallClient> getFunctionCallClient(const std::string& otherHost)
{
    auto it = tlsClients.find(otherHost);
    if (it != tlsClients.end()) {
        return it->second;
    }
    auto c = std::make_shared<FunctionCallClient>(otherHost);
    tlsClients[otherHost] = c;
    return c;
}

void clearFunctionCallClients()
{
    tlsClients.clear();
}

FunctionCallClient::FunctionCallClient(const std::string& hostIn)
  : faabric::transport::MessageEndpointClient(hostIn,
                                              FUNCTION_CALL_ASYNC_PORT,
                                              FUNCTION_CALL_SYNC_PORT)
{}

void FunctionCallClient::sendFlush()
{
    faabric::EmptyRequest req;
    if (faabric::util::isMockMode()) {
        std::lock_guard<std::mutex> lk(mockMutex);
        flushCalls.emplace_back(host, req);
        return;
    }
    faabric::EmptyResponse resp;
    syncSend(FunctionCalls::Flush, &req, &resp);
}

// The destination host name travels with the batch so that one worker process
// can serve several virtual hosts (one per GPU): [u16 len][host][BER]
static std::string wrapWithHost(const std::string& host, const std::string& payload)
{
    std::string out;
    uint16_t n = (uint16_t)host.size();
    out.reserve(2 + host.size() + payload.size());
    out.append((const char*)&n, 2);
    out.append(host);
    out.append(payload);
    return out;
}

void FunctionCallClient::executeFunctions(std::shared_ptr<faabric::BatchExecuteRequest> req)
{
    if (faabric::util::isMockMode()) {
        std::lock_guard<std::mutex> lk(mockMutex);
        batchMessages.emplace_back(host, req);
        return;
    }
    if (!faabric::util::FaultInjector::get().armed()) {
        if (auto* local = faabric::transport::MessageEndpointServer::localServerFor(host, FUNCTION_CALL_ASYNC_PORT, false)) {
            // Same process (a worker serving this - possibly virtual - host):
            // hand the request object to the server's workers as it is.  The
            // executing side stamps and fills its messages, so the caller
            // must not touch the request afterwards (over TCP it would have
            // been serialised; the planner dispatches copies of its own)
            std::string target = host;
            local->getAsyncHandler()->deliverLocalTask([req, target] {
                auto& sch = faabric::scheduler::getScheduler();
                long now = faabric::util::getGlobalClock().epochMillis();
                for (int i = 0; i < req->messages_size(); i++) {
                    auto* m = req->mutable_messages(i);
                    m->set_starttimestamp(now);
                    m->set_executedhost(target.empty() ? sch.getThisHost() : target);
                }
                sch.executeBatch(req);
            });
            return;
        }
    }
    std::string buf = wrapWithHost(host, req->SerializeAsString());
    asyncSend(FunctionCalls::ExecuteFunctions, (const uint8_t*)buf.data(), buf.size());
}

void FunctionCallClient::setMessageResult(std::shared_ptr<faabric::Message> msg)
{
    if (faabric::util::isMockMode()) {
        std::lock_guard<std::mutex> lk(mockMutex);
        messageResults.emplace_back(host, msg);
        return;
    }
    if (!faabric::util::FaultInjector::get().armed() &&
        faabric::transport::MessageEndpointServer::localServerFor(host, FUNCTION_CALL_ASYNC_PORT, false) != nullptr) {
        // The waiter lives in this process: fulfil its promise here (a mutex
        // and a wake-up; nothing to serialise, no worker thread in between)
        faabric::planner::getPlannerClient().setMessageResultLocally(msg, true);
        return;
    }
    asyncSend(FunctionCalls::SetMessageResult, msg.get());
}

// ---------------------------------------------------------------------------
// Server
// ---------------------------------------------------------------------------
FunctionCallServer::FunctionCallServer()
  : faabric::transport::MessageEndpointServer(FUNCTION_CALL_ASYNC_PORT,
                                              FUNCTION_CALL_SYNC_PORT,
                                              FUNCTION_INPROC_LABEL,
                                              faabric::util::getSystemConfig().functionServerThreads)
  , scheduler(getScheduler())
{}

void FunctionCallServer::doAsyncRecv(transport::Message& message)
{
    uint8_t header = message.getMessageCode();
    switch (header) {
        case FunctionCalls::ExecuteFunctions:
            recvExecuteFunctions(message.udata());
            break;
        case FunctionCalls::SetMessageResult:
            recvSetMessageResult(message.udata());
            break;
        default:
            throw std::runtime_error("Unrecognized async call header: " + std::to_string(header));
    }
}

std::string FunctionCallServer::doSyncRecv(transport::Message& message)
{
    uint8_t header = message.getMessageCode();
    if (header == FunctionCalls::Flush) {
        return recvFlush(message.udata());
    }
    throw std::runtime_error("Unrecognized sync call header: " + std::to_string(header));
}

std::string FunctionCallServer::recvFlush(std::span<const uint8_t> buffer)
{
    // Clear out any cached state, executors and the factory's host state
    faabric::state::getGlobalState().forceClearAll(false);
    scheduler.flushLocally();
    // Finished MPI worlds stay registered for late joiners (as in the
    // reference); a flush is where a long-running worker lets go of them
    faabric::mpi::getMpiWorldRegistry().clear();
    return faabric::EmptyResponse().SerializeAsString();
}

void FunctionCallServer::recvExecuteFunctions(std::span<const uint8_t> buffer)
{
    if (buffer.size() < 2) {
        throw std::runtime_error("Malformed ExecuteFunctions payload");
    }
    uint16_t n;
    memcpy(&n, buffer.data(), 2);
    if (buffer.size() < (size_t)2 + n) {
        throw std::runtime_error("Malformed ExecuteFunctions payload");
    }
    std::string targetHost((const char*)buffer.data() + 2, n);
    auto req = std::make_shared<faabric::BatchExecuteRequest>();
    if (!req->ParseFromArray(buffer.data() + 2 + n, (int)(buffer.size() - 2 - n))) {
        throw std::runtime_error("Could not parse batch execute request");
    }
    // This host is now executing the messages: stamp them
    long now = faabric::util::getGlobalClock().epochMillis();
    for (int i = 0; i < req->messages_size(); i++) {
        auto* m = req->mutable_messages(i);
        m->set_starttimestamp(now);
        m->set_executedhost(targetHost.empty() ? scheduler.getThisHost() : targetHost);
    }
    scheduler.executeBatch(req);
}

void FunctionCallServer::recvSetMessageResult(std::span<const uint8_t> buffer)
{
    auto msg = std::make_shared<faabric::Message>();
    if (!msg->ParseFromArray(buffer.data(), (int)buffer.size())) {
        throw std::runtime_error("Could not parse message result");
    }
    faabric::planner::getPlannerClient().setMessageResultLocally(msg, true);
}

} // namespace faabric::scheduler
