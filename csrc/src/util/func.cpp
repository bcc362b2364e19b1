#include <faabric/util/batch.h>
#include <faabric/util/bytes.h>
#include <faabric/util/clock.h>
#include <faabric/util/config.h>
#include <faabric/util/func.h>
#include <faabric/util/gids.h>
#include <faabric/util/logging.h>
#include <faabric/util/string_tools.h>

#include <sstream>

namespace faabric::util {

std::string funcToString(const faabric::Message& msg, bool includeId)
{
    std::string s = msg.user() + "/" + msg.function();
    if (includeId) {
        s += ":" + std::to_string(msg.id());
    }
    return s;
}

std::string funcToString(
  const std::shared_ptr<faabric::BatchExecuteRequest>& req)
{
    return req->user() + "/" + req->function() + ":" +
           std::to_string(req->appid());
}

std::string buildAsyncResponse(const faabric::Message& msg)
{
    if (msg.id() == 0) {
        throw std::runtime_error("Message must have an id before building response");
    }
    return std::to_string(msg.id());
}

std::string resultKeyFromMessageId(unsigned int mid)
{
    return "result_" + std::to_string(mid);
}

std::string statusKeyFromMessageId(unsigned int mid)
{
    return "status_" + std::to_string(mid);
}

unsigned int setMessageId(faabric::Message& msg)
{
    unsigned int id = 0;
    if (msg.id() > 0) {
        id = (unsigned int)msg.id();
    } else {
        id = generateGid();
        msg.set_id((int32_t)id);
    }
    if (msg.appid() == 0) {
        msg.set_appid((int32_t)generateGid());
    }
    // (reference: src/util/func.cpp:102-106 - a message without a timestamp gets one)
    if (msg.starttimestamp() <= 0) {
        msg.set_starttimestamp(getGlobalClock().epochMillis());
    }
    msg.set_resultkey(resultKeyFromMessageId(id));
    msg.set_statuskey(statusKeyFromMessageId(id));
    return id;
}

static void fillMessage(faabric::Message& msg,
                        const std::string& user,
                        const std::string& function)
{
    msg.set_user(user);
    msg.set_function(function);
    setMessageId(msg);
    // Same form as transport::getThisHostAddress(): several workers may share
    // one IP and differ by port offset
    const auto& conf = getSystemConfig();
    msg.set_mainhost(conf.portOffset == 0 ? conf.endpointHost
                                          : conf.endpointHost + ":" + std::to_string(conf.portOffset));
}

std::shared_ptr<faabric::Message> messageFactoryShared(
  const std::string& user,
  const std::string& function)
{
    auto msg = std::make_shared<faabric::Message>();
    fillMessage(*msg, user, function);
    return msg;
}

faabric::Message messageFactory(const std::string& user,
                                const std::string& function)
{
    faabric::Message msg;
    fillMessage(msg, user, function);
    return msg;
}

std::vector<uint8_t> messageToBytes(const faabric::Message& msg)
{
    std::string s = msg.SerializeAsString();
    return std::vector<uint8_t>(s.begin(), s.end());
}

std::vector<std::string> getArgvForMessage(const faabric::Message& msg)
{
    // Function name first, then whitespace-separated cmdline
    std::vector<std::string> argv = { "function.wasm" };
    std::istringstream is(msg.cmdline());
    std::string tok;
    while (is >> tok) {
        argv.push_back(tok);
    }
    return argv;
}

std::string getMainThreadSnapshotKey(const faabric::Message& msg)
{
    if (msg.appid() == 0) {
        throw std::runtime_error("Message must have an app id to get snapshot key");
    }
    return funcToString(msg, false) + "_" + std::to_string(msg.appid());
}

// ------------------------------------------------------------------ batch ---
std::shared_ptr<faabric::BatchExecuteRequest> batchExecFactory()
{
    auto req = std::make_shared<faabric::BatchExecuteRequest>();
    req->set_appid((int32_t)generateGid());
    return req;
}

std::shared_ptr<faabric::BatchExecuteRequest> batchExecFactory(
  const std::string& user,
  const std::string& function,
  int count)
{
    auto req = batchExecFactory();
    req->set_user(user);
    req->set_function(function);
    // All messages of a batch share the app id and are indexed within it
    for (int i = 0; i < count; i++) {
        faabric::Message* m = req->add_messages();
        *m = messageFactory(user, function);
        m->set_appid(req->appid());
        m->set_appidx(i);
    }
    return req;
}

bool isBatchExecRequestValid(std::shared_ptr<faabric::BatchExecuteRequest> ber)
{
    if (ber == nullptr) {
        SPDLOG_ERROR("Invalid BER: null");
        return false;
    }
    // never initialised: no messages and no app id
    if (ber->messages_size() <= 0 && ber->appid() == 0) {
        SPDLOG_ERROR("Invalid BER: zero messages");
        return false;
    }
    if (ber->user().empty() || ber->function().empty()) {
        SPDLOG_ERROR("Invalid BER: empty user or function");
        return false;
    }
    // Every message shares the request's user and app id.  Function names may
    // differ (a request may carry calls chained by name) but not be empty
    // (reference: src/util/batch.cpp:58-78)
    for (int i = 0; i < ber->messages_size(); i++) {
        const auto& m = ber->messages(i);
        if (m.user() != ber->user() || m.function().empty() || m.appid() != ber->appid()) {
            SPDLOG_ERROR("Invalid BER: message {} inconsistent with request", i);
            return false;
        }
    }
    return true;
}

void updateBatchExecAppId(std::shared_ptr<faabric::BatchExecuteRequest> ber,
                          int newAppId)
{
    ber->set_appid(newAppId);
    for (int i = 0; i < ber->messages_size(); i++) {
        ber->mutable_messages(i)->set_appid(newAppId);
    }
    // (An empty request under construction is fine)
    if (ber->messages_size() > 0 && !isBatchExecRequestValid(ber)) {
        throw std::runtime_error("Invalid BER after updating app id");
    }
}

void updateBatchExecGroupId(std::shared_ptr<faabric::BatchExecuteRequest> ber,
                            int newGroupId)
{
    ber->set_groupid(newGroupId);
    for (int i = 0; i < ber->messages_size(); i++) {
        ber->mutable_messages(i)->set_groupid(newGroupId);
    }
    if (ber->messages_size() > 0 && !isBatchExecRequestValid(ber)) {
        throw std::runtime_error("Invalid BER after updating group id");
    }
}

std::shared_ptr<faabric::BatchExecuteRequestStatus> batchExecStatusFactory(
  int32_t appId)
{
    auto st = std::make_shared<faabric::BatchExecuteRequestStatus>();
    st->set_appid(appId);
    st->set_finished(false);
    return st;
}

std::shared_ptr<faabric::BatchExecuteRequestStatus> batchExecStatusFactory(
  std::shared_ptr<faabric::BatchExecuteRequest> ber)
{
    auto st = batchExecStatusFactory(ber->appid());
    st->set_expectednummessages(ber->messages_size());
    return st;
}

} // namespace faabric::util
