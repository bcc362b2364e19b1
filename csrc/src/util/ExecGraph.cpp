#include <faabric/planner/PlannerClient.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/func.h>
#include <faabric/util/json.h>
#include <faabric/util/logging.h>

#include <functional>

namespace faabric::util {

// Results should already be there: a short timeout is enough
static constexpr int EXEC_GRAPH_TIMEOUT_MS = 1000;

ExecGraphNode getFunctionExecGraphNode(int appId, int msgId)
{
    auto& client = faabric::planner::getPlannerClient();
    faabric::Message result = client.getMessageResult(appId, msgId, EXEC_GRAPH_TIMEOUT_MS);
    if (result.type() == faabric::Message::EMPTY) {
        throw ExecGraphNodeNotFoundException("Exec graph node not found: " + std::to_string(msgId));
    }
    ExecGraphNode node;
    node.msg = result;
    for (unsigned int childId : getChainedFunctions(result)) {
        node.children.push_back(getFunctionExecGraphNode(appId, (int)childId));
    }
    return node;
}

ExecGraph getFunctionExecGraph(const faabric::Message& msg)
{
    try {
        return ExecGraph{ getFunctionExecGraphNode(msg.appid(), msg.id()) };
    } catch (ExecGraphNodeNotFoundException& e) {
        SPDLOG_ERROR("Could not build exec graph for app {} msg {}: {}", msg.appid(), msg.id(), e.what());
        return ExecGraph{};
    }
}

void logChainedFunction(faabric::Message& parentMessage, const faabric::Message& chainedMessage)
{
    parentMessage.add_chainedmsgids(chainedMessage.id());
}

std::set<unsigned int> getChainedFunctions(const faabric::Message& msg)
{
    std::set<unsigned int> out;
    for (auto id : msg.chainedmsgids()) {
        out.insert((unsigned int)id);
    }
    return out;
}

static void walk(const ExecGraphNode& node, const std::function<void(const ExecGraphNode&)>& fn)
{
    fn(node);
    for (const auto& c : node.children) {
        walk(c, fn);
    }
}

int countExecGraphNodes(const ExecGraph& graph)
{
    int n = 0;
    walk(graph.rootNode, [&](const ExecGraphNode&) { n++; });
    return n;
}

std::set<std::string> getExecGraphHosts(const ExecGraph& graph)
{
    std::set<std::string> hosts;
    walk(graph.rootNode, [&](const ExecGraphNode& n) { hosts.insert(n.msg.executedhost()); });
    return hosts;
}

std::vector<std::string> getMpiRankHostsFromExecGraph(const ExecGraph& graph)
{
    int worldSize = graph.rootNode.msg.mpiworldsize();
    std::vector<std::string> hosts((size_t)std::max(worldSize, 0));
    walk(graph.rootNode, [&](const ExecGraphNode& n) {
        int rank = n.msg.mpirank();
        if (rank >= 0 && rank < (int)hosts.size()) {
            hosts[(size_t)rank] = n.msg.executedhost();
        }
    });
    return hosts;
}

std::pair<std::vector<std::string>, std::vector<std::string>> getMigratedMpiRankHostsFromExecGraph(
  const ExecGraph& graph)
{
    // A migrated rank shows up twice: once with a MIGRATED return value on
    // the origin host, and once as a chained child on the destination
    int worldSize = graph.rootNode.msg.mpiworldsize();
    std::vector<std::string> before((size_t)std::max(worldSize, 0));
    std::vector<std::string> after((size_t)std::max(worldSize, 0));
    std::function<void(const ExecGraphNode&, bool)> rec = [&](const ExecGraphNode& n, bool parentMigrated) {
        int rank = n.msg.mpirank();
        bool migrated = n.msg.returnvalue() == MIGRATED_FUNCTION_RETURN_VALUE;
        if (rank >= 0 && rank < worldSize) {
            if (migrated || before[(size_t)rank].empty()) {
                before[(size_t)rank] = n.msg.executedhost();
            }
            if (!migrated) {
                after[(size_t)rank] = n.msg.executedhost();
            }
        }
        for (const auto& c : n.children) {
            rec(c, migrated);
        }
    };
    rec(graph.rootNode, false);
    return { before, after };
}

static void nodeToJson(const ExecGraphNode& node, faabric::proto::JsonWriter& w)
{
    w.beginObject();
    w.key("msg");
    w.raw(messageToJson(node.msg));
    if (!node.children.empty()) {
        w.key("chained");
        w.beginArray();
        for (const auto& c : node.children) {
            nodeToJson(c, w);
        }
        w.endArray();
    }
    w.endObject();
}

std::string execNodeToJson(const ExecGraphNode& node)
{
    faabric::proto::JsonWriter w;
    nodeToJson(node, w);
    return w.str();
}

std::string execGraphToJson(const ExecGraph& graph)
{
    faabric::proto::JsonWriter w;
    w.beginObject();
    w.key("root");
    nodeToJson(graph.rootNode, w);
    w.endObject();
    return w.str();
}

void addDetail(faabric::Message& msg, const std::string& key, const std::string& value)
{
    if (!msg.recordexecgraph()) {
        return;
    }
    (*msg.mutable_execgraphdetails())[key] = value;
}

void incrementCounter(faabric::Message& msg, const std::string& key, int valueToIncrement)
{
    if (!msg.recordexecgraph()) {
        return;
    }
    (*msg.mutable_intexecgraphdetails())[key] += valueToIncrement;
}

}
