// Smaller components (strategy: reference tests/test/util/test_concurrent_map,
// test_files, test_hwloc; executor/test_executor_context; transport/
// test_tcp_sockets, test_message; proto/test_proto; snapshot/
// test_snapshot_registry; scheduler/test_function_client_server;
// runner/test_main; mpi/test_mpi_exec_graph, test_multiple_mpi_worlds)
#include "fixtures.h"

#include <faabric/executor/ExecutorContext.h>
#include <faabric/mpi/MpiWorldRegistry.h>
#include <faabric/mpi/mpi.h>
#include <faabric/runner/FaabricMain.h>
#include <faabric/scheduler/FunctionCallClient.h>
#include <faabric/snapshot/SnapshotRegistry.h>
#include <faabric/transport/tcp/Socket.h>
#include <faabric/util/ExecGraph.h>
#include <faabric/util/compare.h>
#include <faabric/util/concurrent_map.h>
#include <faabric/util/files.h>
#include <faabric/util/hwloc.h>

#include <filesystem>
#include <thread>
#include <unistd.h>

using namespace tests;

TEST_CASE("concurrent map operations", "[util]")
{
    faabric::util::ConcurrentMap<int, std::shared_ptr<int>> map;
    REQUIRE(map.isEmpty());
    REQUIRE(map.tryEmplace(1, std::make_shared<int>(10)));
    REQUIRE(!map.tryEmplace(1, std::make_shared<int>(11)));
    auto [inserted, value] = map.tryEmplaceShared(2, std::make_shared<int>(20));
    REQUIRE(inserted);
    REQUIRE_EQ(*value, 20);
    auto [again, same] = map.tryEmplaceShared(2, std::make_shared<int>(21));
    REQUIRE(!again);
    REQUIRE_EQ(*same, 20);
    bool wasInserted = true;
    map.tryEmplaceThenMutate(
      2, [&](bool ins, std::shared_ptr<int>& v) {
          wasInserted = ins;
          *v += 5;
      },
      std::make_shared<int>(0));
    REQUIRE(!wasInserted);
    REQUIRE_EQ(**map.get(2), 25);
    map.insertOrAssign(3, std::make_shared<int>(30));
    REQUIRE(map.contains(3));
    REQUIRE(!map.get(99).has_value());
    int seen = 0;
    REQUIRE(map.inspect(3, [&](const std::shared_ptr<int>& v) { seen = *v; }));
    REQUIRE_EQ(seen, 30);
    REQUIRE_EQ(map.eraseIf([](const int& k, const std::shared_ptr<int>& v) { return *v >= 25; }), 2u);
    REQUIRE_EQ(map.size(), 1u);
    REQUIRE(map.erase(1));
    REQUIRE(!map.erase(1));

    // Concurrent emplaces of distinct keys all land
    faabric::util::ConcurrentMap<int, int> counts;
    std::vector<std::thread> ts;
    for (int t = 0; t < 8; t++) {
        ts.emplace_back([&, t] {
            for (int i = 0; i < 500; i++) {
                counts.tryEmplace(t * 1000 + i, i);
            }
        });
    }
    for (auto& t : ts) {
        t.join();
    }
    REQUIRE_EQ(counts.size(), 4000u);
}

TEST_CASE("files, compare and cpu pinning helpers", "[util]")
{
    std::string path = "/tmp/faabric_b200_test_" + std::to_string(getpid());
    std::vector<uint8_t> bytes = { 0x00, 0x61, 0x73, 0x6d, 1, 0, 0, 0 };
    faabric::util::writeBytesToFile(path, bytes);
    REQUIRE(faabric::util::readFileToBytes(path) == bytes);
    REQUIRE_EQ(faabric::util::readFileToString(path).size(), bytes.size());
    REQUIRE(faabric::util::isWasm(bytes));
    REQUIRE(!faabric::util::isWasm({ 1, 2, 3, 4 }));
    ::unlink(path.c_str());
    REQUIRE_THROWS(faabric::util::readFileToBytes("/nonexistent/file"));

    int a[3] = { 1, 2, 3 }, b[3] = { 1, 2, 3 }, c[3] = { 1, 2, 4 };
    REQUIRE(faabric::util::compareArrays(a, b, 3));
    REQUIRE(!faabric::util::compareArrays(a, c, 3));

    int before = faabric::util::getNumFreeCpus();
    REQUIRE(before > 0);
    {
        std::vector<std::unique_ptr<faabric::util::FaabricCpuSet>> pins;
        std::thread t([&] { pins.push_back(faabric::util::pinThreadToFreeCpu(pthread_self())); });
        t.join();
        REQUIRE_EQ(faabric::util::getNumFreeCpus(), before - 1);
        REQUIRE(pins[0]->getCpuIdx() >= 0);
    }
    // Released when the handle goes away
    REQUIRE_EQ(faabric::util::getNumFreeCpus(), before);
}

TEST_CASE("executor context is per thread", "[executor]")
{
    using faabric::executor::ExecutorContext;
    REQUIRE(!ExecutorContext::isSet());
    REQUIRE_THROWS(ExecutorContext::get());
    auto req = faabric::util::batchExecFactory("demo", "ctx", 3);
    ExecutorContext::set(nullptr, req, 2);
    REQUIRE(ExecutorContext::isSet());
    REQUIRE_EQ(ExecutorContext::get()->getMsgIdx(), 2);
    REQUIRE_EQ(ExecutorContext::get()->getMsg().id(), req->messages(2).id());
    std::thread other([&] {
        if (ExecutorContext::isSet()) {
            fbtest::fail(__FILE__, __LINE__, "context leaked into another thread");
        }
    });
    other.join();
    ExecutorContext::unset();
    REQUIRE(!ExecutorContext::isSet());
}

TEST_CASE("raw tcp sockets", "[transport]")
{
    using namespace faabric::transport::tcp;
    int port = 9733;
    RecvSocket server(port);
    server.listen();
    std::vector<uint8_t> payload(200000);
    for (size_t i = 0; i < payload.size(); i++) {
        payload[i] = (uint8_t)(i * 7);
    }
    std::thread client([&] {
        SendSocket s("127.0.0.1", port);
        s.dial();
        uint32_t n = (uint32_t)payload.size();
        s.sendOne((const uint8_t*)&n, sizeof(n));
        s.sendOne(payload.data(), payload.size());
    });
    int conn = server.accept(5000);
    REQUIRE(conn >= 0);
    uint32_t n = 0;
    server.recvOne(conn, (uint8_t*)&n, sizeof(n));
    REQUIRE_EQ(n, (uint32_t)payload.size());
    std::vector<uint8_t> got(n);
    server.recvOne(conn, got.data(), n);
    client.join();
    REQUIRE(got == payload);
    // Nobody listening: dial gives up
    SendSocket nobody("127.0.0.1", 9734);
    REQUIRE_THROWS(nobody.dial(2, 10));
}

TEST_CASE("raw tcp sockets: timeouts, closed peers, moved handles", "[transport]")
{
    using namespace faabric::transport::tcp;
    int port = 9735;
    RecvSocket server(port);
    server.listen();
    REQUIRE_EQ(server.getPort(), port);
    // nobody connects: accept gives up after the timeout
    auto t0 = std::chrono::steady_clock::now();
    REQUIRE_THROWS(server.accept(50));
    auto waited = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    REQUIRE(waited >= 40 && waited < 2000);

    // the port is taken: a second listener is refused
    RecvSocket clash(port);
    REQUIRE_THROWS(clash.listen());

    // a peer that dies mid-message surfaces as an error, not a hang
    std::thread client([&] {
        SendSocket s("127.0.0.1", port);
        s.dial();
        uint8_t half[10] = { 1, 2, 3, 4, 5, 6, 7, 8, 9, 10 };
        s.sendOne(half, sizeof(half));
        // destructor closes the connection with 10 of 64 bytes sent
    });
    int conn = server.accept(5000);
    client.join();
    std::vector<uint8_t> buf(64);
    REQUIRE_THROWS(server.recvOne(conn, buf.data(), buf.size()));

    // sockets are movable handles
    Socket a;
    int fd = a.get();
    REQUIRE(fd >= 0);
    Socket b(std::move(a));
    REQUIRE_EQ(b.get(), fd);
    REQUIRE(a.get() < 0);
    Socket c;
    c = std::move(b);
    REQUIRE_EQ(c.get(), fd);
    c.close();
    REQUIRE(c.get() < 0);
    // sending on an unconnected socket fails cleanly
    SendSocket idle("127.0.0.1", port + 1);
    uint8_t byte = 0;
    REQUIRE_THROWS(idle.sendOne(&byte, 1));
}

TEST_CASE("transport message header and wire format", "[transport][proto]")
{
    uint8_t hdr[HEADER_MSG_SIZE];
    faabric::transport::Message::writeHeader(hdr, 42, 123456789012ull, -7);
    uint8_t code;
    uint64_t size;
    int32_t seq;
    faabric::transport::Message::readHeader(hdr, code, size, seq);
    REQUIRE_EQ((int)code, 42);
    REQUIRE_EQ(size, 123456789012ull);
    REQUIRE_EQ(seq, -7);

    // Protobuf wire compatibility: field 1 varint, field 6 string
    faabric::Message m;
    m.set_id(150);
    m.set_user("ab");
    std::string wire = m.SerializeAsString();
    std::string expected = { 0x08, (char)0x96, 0x01, 0x32, 0x02, 'a', 'b' };
    REQUIRE_EQ(wire, expected);
    // Unknown fields are skipped, negative ints are 10-byte varints
    faabric::Message neg;
    neg.set_returnvalue(-1);
    REQUIRE_EQ(neg.SerializeAsString().size(), 11u);
    faabric::Message parsed;
    std::string withUnknown = wire + std::string({ (char)0xf8, 0x7f, 0x05 }); // field 2047 varint 5
    REQUIRE(parsed.ParseFromString(withUnknown));
    REQUIRE_EQ(parsed.id(), 150);
    REQUIRE(!parsed.ParseFromString(std::string({ 0x0a, 0x7f }))); // truncated length-delimited field
}

TEST_CASE("snapshot registry", "[snapshot]")
{
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    reg.clear();
    REQUIRE_EQ(reg.getSnapshotCount(), 0u);
    auto a = std::make_shared<faabric::util::SnapshotData>(1024);
    auto b = std::make_shared<faabric::util::SnapshotData>(2048);
    reg.registerSnapshot("a", a);
    reg.registerSnapshot("b", b);
    REQUIRE_EQ(reg.getSnapshotCount(), 2u);
    REQUIRE(reg.getSnapshot("b").get() == b.get());
    // Re-registering replaces
    auto a2 = std::make_shared<faabric::util::SnapshotData>(4096);
    reg.registerSnapshot("a", a2);
    REQUIRE_EQ(reg.getSnapshot("a")->getSize(), 4096u);
    reg.deleteSnapshot("a");
    REQUIRE(!reg.snapshotExists("a"));
    REQUIRE_THROWS(reg.getSnapshot("a"));
    reg.clear();
}

TEST_CASE("function call client records calls in mock mode", "[scheduler]")
{
    faabric::util::setMockMode(true);
    faabric::scheduler::clearMockRequests();
    auto req = faabric::util::batchExecFactory("demo", "mock", 2);
    faabric::scheduler::FunctionCallClient cli("somewhere");
    cli.executeFunctions(req);
    cli.sendFlush();
    auto res = std::make_shared<faabric::Message>(req->messages(0));
    cli.setMessageResult(res);
    auto batches = faabric::scheduler::getBatchRequests();
    REQUIRE_EQ(batches.size(), 1u);
    REQUIRE_EQ(batches[0].first, std::string("somewhere"));
    REQUIRE_EQ(batches[0].second->messages_size(), 2);
    REQUIRE_EQ(faabric::scheduler::getFlushCalls().size(), 1u);
    REQUIRE_EQ(faabric::scheduler::getMessageResults().size(), 1u);
    faabric::scheduler::clearMockRequests();
    REQUIRE_EQ(faabric::scheduler::getBatchRequests().size(), 0u);
    faabric::util::setMockMode(false);
}

TEST_CASE("runner boots and shuts a worker down", "[runner]")
{
    // Planner in-process, then a full worker through FaabricMain
    faabric::planner::PlannerServer plannerServer;
    plannerServer.start();
    faabric::planner::getPlanner().reset();
    {
        faabric::runner::FaabricMain m(std::make_shared<TestExecutorFactory>());
        m.startBackground();
        auto hosts = faabric::planner::getPlannerClient().getAvailableHosts();
        REQUIRE_EQ(hosts.size(), 1u);
        REQUIRE(hosts[0].slots() > 0);
        auto req = faabric::util::batchExecFactory("demo", "echo", 1);
        req->mutable_messages(0)->set_inputdata("via runner");
        faabric::planner::getPlannerClient().callFunctions(req);
        auto res = faabric::planner::getPlannerClient().getMessageResult(req->messages(0), 5000);
        REQUIRE_EQ(res.outputdata(), std::string("via runner"));
        m.shutdown();
        REQUIRE_EQ(faabric::planner::getPlannerClient().getAvailableHosts().size(), 0u);
    }
    faabric::planner::getPlanner().reset();
    plannerServer.stop();
    faabric::scheduler::getScheduler().reset();
}

TEST_CASE("local cluster: one process serves a planner and per-GPU virtual hosts", "[runner]")
{
    registerTestFunction("demo", "where", [](auto*, int, int idx, auto req) {
        auto& m = *req->mutable_messages(idx);
        m.set_outputdata(m.executedhost());
        return 0;
    });
    {
        faabric::runner::LocalCluster cluster(std::make_shared<TestExecutorFactory>(), 4, 2);
        REQUIRE_EQ(cluster.hosts().size(), (size_t)4);
        auto hosts = faabric::planner::getPlannerClient().getAvailableHosts();
        REQUIRE_EQ(hosts.size(), (size_t)4);
        std::set<std::string> names;
        for (auto& h : hosts) {
            names.insert(h.ip());
            REQUIRE_EQ(h.slots(), 2);
        }
        REQUIRE(names == std::set<std::string>(cluster.hosts().begin(), cluster.hosts().end()));
        // eight functions fill every virtual host
        auto req = faabric::util::batchExecFactory("demo", "where", 8);
        auto decision = faabric::planner::getPlannerClient().callFunctions(req);
        REQUIRE_EQ(decision.nFunctions, 8);
        auto status = cluster.awaitBatch(req, 20000);
        REQUIRE_EQ(status->messageresults_size(), 8);
        std::map<std::string, int> perHost;
        for (auto& m : status->messageresults()) {
            REQUIRE_EQ(m.returnvalue(), 0);
            perHost[m.executedhost()]++;
        }
        REQUIRE_EQ(perHost.size(), (size_t)4);
        for (auto& [h, n] : perHost) {
            REQUIRE_EQ(n, 2);
        }
        // a batch nobody submitted is not in flight: no results, no waiting
        auto ghost = faabric::util::batchExecFactory("demo", "where", 1);
        REQUIRE_EQ(cluster.awaitBatch(ghost, 100)->messageresults_size(), 0);
    }
    // a plain cluster: just this host
    {
        faabric::runner::LocalCluster plain(std::make_shared<TestExecutorFactory>(), 0, 3);
        REQUIRE_EQ(plain.hosts().size(), (size_t)1);
        auto hosts = faabric::planner::getPlannerClient().getAvailableHosts();
        REQUIRE_EQ(hosts.size(), (size_t)1);
        REQUIRE_EQ(hosts[0].slots(), 3);
    }
    faabric::scheduler::getScheduler().reset();
}

TEST_CASE("runner checkpoints snapshots across a worker restart", "[runner][checkpoint]")
{
    const std::string dir = "/tmp/fb_runner_ckpt_" + std::to_string(getpid());
    std::filesystem::remove_all(dir);
    auto& conf = faabric::util::getSystemConfig();
    conf.checkpointDir = dir;
    faabric::planner::PlannerServer plannerServer;
    plannerServer.start();
    faabric::planner::getPlanner().reset();
    auto& reg = faabric::snapshot::getSnapshotRegistry();
    reg.clear();
    std::vector<uint8_t> image(10000, 42);
    {
        faabric::runner::FaabricMain m(std::make_shared<TestExecutorFactory>());
        m.startBackground();
        reg.registerSnapshot("migration_9001", std::make_shared<faabric::util::SnapshotData>(image));
        m.shutdown();
    }
    // "restart": the registry is empty until the next worker boots
    reg.clear();
    {
        faabric::runner::FaabricMain m(std::make_shared<TestExecutorFactory>());
        m.startBackground();
        REQUIRE(reg.snapshotExists("migration_9001"));
        REQUIRE(reg.getSnapshot("migration_9001")->getDataCopy() == image);
        m.shutdown();
    }
    conf.checkpointDir.clear();
    reg.clear();
    faabric::planner::getPlanner().reset();
    plannerServer.stop();
    faabric::scheduler::getScheduler().reset();
    std::filesystem::remove_all(dir);
}

TEST_CASE("mpi: exec graph counts messages, two worlds run side by side", "[mpi]")
{
    ClusterFixture f(8);
    registerTestFunction("mpi", "graph", [&](auto*, int, int idx, auto req) {
        MPI_Init(nullptr, nullptr);
        int rank, size;
        MPI_Comm_rank(MPI_COMM_WORLD, &rank);
        MPI_Comm_size(MPI_COMM_WORLD, &size);
        int v = rank;
        if (rank == 0) {
            for (int i = 0; i < 3; i++) {
                MPI_Send(&v, 1, MPI_INT, 1, 0, MPI_COMM_WORLD);
            }
        } else if (rank == 1) {
            for (int i = 0; i < 3; i++) {
                MPI_Recv(&v, 1, MPI_INT, 0, 0, MPI_COMM_WORLD, MPI_STATUS_IGNORE);
            }
        }
        int sum = 0;
        MPI_Allreduce(&rank, &sum, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
        MPI_Finalize();
        return sum == size * (size - 1) / 2 ? 0 : 1;
    });
    // Two independent worlds of different sizes at the same time
    auto reqA = faabric::util::batchExecFactory("mpi", "graph", 1);
    reqA->mutable_messages(0)->set_ismpi(true);
    reqA->mutable_messages(0)->set_mpiworldsize(3);
    reqA->mutable_messages(0)->set_recordexecgraph(true);
    auto reqB = faabric::util::batchExecFactory("mpi", "graph", 1);
    reqB->mutable_messages(0)->set_ismpi(true);
    reqB->mutable_messages(0)->set_mpiworldsize(4);
    f.plannerCli.callFunctions(reqA);
    f.plannerCli.callFunctions(reqB);
    auto stA = f.awaitBatch(reqA);
    auto stB = f.awaitBatch(reqB);
    REQUIRE_EQ(stA->messageresults_size(), 3);
    REQUIRE_EQ(stB->messageresults_size(), 4);
    std::set<int> worlds;
    for (auto* st : { stA.get(), stB.get() }) {
        for (auto& m : st->messageresults()) {
            REQUIRE_EQ(m.returnvalue(), 0);
            worlds.insert(m.mpiworldid());
        }
    }
    REQUIRE_EQ(worlds.size(), 2u);
    // Rank 0 of world A recorded its three normal messages to rank 1
    for (auto& m : stA->messageresults()) {
        if (m.mpirank() == 0) {
            std::string key = std::string(MPI_MSGTYPE_COUNT_PREFIX) + "-" + std::to_string((int)faabric::mpi::MpiMessageType::NORMAL) + "-1";
            REQUIRE(m.intexecgraphdetails().count(key) == 1);
            REQUIRE_EQ(m.intexecgraphdetails().at(key), 3);
            REQUIRE(m.intexecgraphdetails().at(std::string(MPI_MSG_COUNT_PREFIX) + "-1") >= 3);
            // ...and the chained ranks show up in its exec graph
            auto graph = faabric::util::getFunctionExecGraph(m);
            REQUIRE_EQ(faabric::util::countExecGraphNodes(graph), 3);
            auto hosts = faabric::util::getMpiRankHostsFromExecGraph(graph);
            REQUIRE_EQ(hosts.size(), 3u);
        }
    }
    faabric::mpi::getMpiWorldRegistry().clear();
}

// ---------------------------------------------------------------------------
// Communicator tuning file (no device needed)
// ---------------------------------------------------------------------------
#include <faabric/device/communicator.h>

TEST_CASE("device: tuning file parses, serialises and applies", "[device][config]")
{
    using faabric::device::CommConfig;
    using faabric::device::CommTuning;
    const std::string text = "# measured on 8 GPUs\n"
                             "set oneShotMaxBytes 131072\n"
                             "set  threads   256   # trailing comment\n"
                             "\n"
                             "allreduce 18446744073709551615 nvls\n"
                             "allreduce 4096 ll\n"
                             "allreduce 1048576 twoshot\n";
    CommTuning t = CommTuning::parse(text);
    REQUIRE_EQ(t.settings.size(), (size_t)2);
    REQUIRE_EQ(t.allReduceTable.size(), (size_t)3);
    // rows come back sorted by size
    REQUIRE_EQ(t.allReduceTable[0].first, (uint64_t)4096);
    REQUIRE_EQ(t.allReduceTable[0].second, (int)FB_ALGO_LL);
    REQUIRE_EQ(t.allReduceTable[1].second, (int)FB_ALGO_TWOSHOT);
    REQUIRE_EQ(t.allReduceTable[2].first, UINT64_MAX);
    REQUIRE_EQ(t.allReduceTable[2].second, (int)FB_ALGO_NVLS);

    CommConfig cfg;
    const size_t llBefore = cfg.llMaxBytes;
    t.applyTo(cfg);
    REQUIRE_EQ(cfg.oneShotMaxBytes, (size_t)131072);
    REQUIRE_EQ(cfg.threads, 256);
    REQUIRE_EQ(cfg.llMaxBytes, llBefore);

    // serialise -> parse is a fixed point
    CommTuning again = CommTuning::parse(t.serialise());
    REQUIRE(again.allReduceTable == t.allReduceTable);
    REQUIRE(again.settings == t.settings);
    REQUIRE_EQ(again.serialise(), t.serialise());

    // file round trip
    std::string path = "/tmp/fb_tuning_" + std::to_string(getpid()) + ".txt";
    faabric::util::writeBytesToFile(path, faabric::util::stringToBytes(t.serialise()));
    CommTuning fromFile;
    REQUIRE(CommTuning::loadFile(path, fromFile));
    REQUIRE(fromFile.allReduceTable == t.allReduceTable);
    ::unlink(path.c_str());
    REQUIRE(!CommTuning::loadFile(path, fromFile));

    // malformed input names the line
    for (const char* bad : { "allreduce 4096 warp9\n", "allreduce lots ll\n", "set nope 1\n", "frobnicate\n", "\nset threads\n", "allreduce 1 auto\n" }) {
        bool threw = false;
        try {
            CommTuning::parse(bad);
        } catch (const std::runtime_error& e) {
            threw = std::string(e.what()).find("line") != std::string::npos;
        }
        REQUIRE(threw);
    }
    REQUIRE(CommTuning::parse("").empty());
}

// The reference's runner case, section by section
// (reference: tests/test/runner/test_main.cpp:29-72)
namespace {
void mainRunnerCase(bool makeCalls)
{
    faabric::planner::PlannerServer plannerServer;
    plannerServer.start();
    faabric::planner::getPlanner().reset();
    {
        faabric::runner::FaabricMain m(std::make_shared<TestExecutorFactory>());
        m.startBackground();
        if (makeCalls) {
            auto req = faabric::util::batchExecFactory("demo", "echo", 4);
            // (waiter and executors share an address space: keep ids, not the request)
            const int appId = req->appid();
            std::vector<int> msgIds;
            for (int i = 0; i < 4; i++) {
                req->mutable_messages(i)->set_inputdata("call " + std::to_string(req->messages(i).id()));
                msgIds.push_back(req->messages(i).id());
            }
            faabric::planner::getPlannerClient().callFunctions(req);
            for (int id : msgIds) {
                auto res = faabric::planner::getPlannerClient().getMessageResult(appId, id, 5000);
                REQUIRE_EQ(res.returnvalue(), 0);
                REQUIRE_EQ(res.outputdata(), "call " + std::to_string(id));
            }
        }
        m.shutdown();
    }
    faabric::planner::getPlanner().reset();
    plannerServer.stop();
    faabric::scheduler::getScheduler().reset();
}
}

TEST_CASE("runner case: started in the background and shut down without work", "[runner][cases]")
{
    mainRunnerCase(false);
}

TEST_CASE("runner case: four calls through the planner come back from the worker", "[runner][cases]")
{
    mainRunnerCase(true);
}
