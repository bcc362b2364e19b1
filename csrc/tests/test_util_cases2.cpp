// More util helpers, one case per case of the reference's suites
// (reference: tests/test/util/test_batch.cpp, test_tokens.cpp,
// test_environment.cpp, test_config.cpp, test_latch.cpp, test_barrier.cpp,
// test_locks.cpp, test_random.cpp, test_delta.cpp, test_gids.cpp,
// test_hwloc.cpp, test_json.cpp, test_files.cpp, test_network.cpp)
#include "harness.h"

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/barrier.h>
#include <faabric/util/batch.h>
#include <faabric/util/config.h>
#include <faabric/util/delta.h>
#include <faabric/util/environment.h>
#include <faabric/util/files.h>
#include <faabric/util/func.h>
#include <faabric/util/gids.h>
#include <faabric/util/hwloc.h>
#include <faabric/util/json.h>
#include <faabric/util/latch.h>
#include <faabric/util/locks.h>
#include <faabric/util/network.h>
#include <faabric/util/queue.h>
#include <faabric/util/random.h>
#include <faabric/util/testing.h>

#include <set>
#include <thread>
#include <unordered_set>

using namespace faabric::util;

// ---- batch ----------------------------------------------------------------
TEST_CASE("batch case: the factory gives every message the request's app id, user and function", "[util][batch][cases]")
{
    auto req = batchExecFactory("demo", "echo", 4);
    REQUIRE_EQ(req->messages_size(), 4);
    REQUIRE(req->appid() > 0);
    for (const auto& m : req->messages()) {
        REQUIRE_EQ(m.appid(), req->appid());
        REQUIRE_EQ(m.user(), std::string("demo"));
        REQUIRE_EQ(m.function(), std::string("echo"));
    }
    // ids differ
    std::set<int> ids;
    for (const auto& m : req->messages()) {
        ids.insert(m.id());
    }
    REQUIRE_EQ(ids.size(), 4u);
}

TEST_CASE("batch case: which requests are well-formed", "[util][batch][cases]")
{
    auto fresh = [] { return batchExecFactory("demo", "echo", 4); };
    REQUIRE(isBatchExecRequestValid(fresh()));
    REQUIRE(!isBatchExecRequestValid(nullptr));
    REQUIRE(!isBatchExecRequestValid(std::make_shared<faabric::BatchExecuteRequest>()));
    auto appMismatch = fresh();
    appMismatch->mutable_messages(1)->set_appid(1337);
    REQUIRE(!isBatchExecRequestValid(appMismatch));
    auto emptyUser = fresh();
    emptyUser->mutable_messages(0)->set_user("");
    REQUIRE(!isBatchExecRequestValid(emptyUser));
    auto emptyFunction = fresh();
    emptyFunction->mutable_messages(0)->set_function("");
    REQUIRE(!isBatchExecRequestValid(emptyFunction));
    auto userMismatch = fresh();
    userMismatch->mutable_messages(1)->set_user("foo");
    REQUIRE(!isBatchExecRequestValid(userMismatch));
    // another function name is fine: calls chained by name
    auto functionMismatch = fresh();
    functionMismatch->mutable_messages(1)->set_function("foo");
    REQUIRE(isBatchExecRequestValid(functionMismatch));
}

TEST_CASE("batch case: updating the app id of a request keeps it well-formed", "[util][batch][cases]")
{
    auto req = batchExecFactory("demo", "echo", 4);
    updateBatchExecAppId(req, 1337);
    REQUIRE(isBatchExecRequestValid(req));
    REQUIRE_EQ(req->appid(), 1337);
    for (const auto& m : req->messages()) {
        REQUIRE_EQ(m.appid(), 1337);
    }
}

TEST_CASE("batch case: updating the group id of a request reaches every message", "[util][batch][cases]")
{
    auto req = batchExecFactory("demo", "echo", 4);
    updateBatchExecGroupId(req, 1337);
    REQUIRE(isBatchExecRequestValid(req));
    REQUIRE_EQ(req->groupid(), 1337);
    for (const auto& m : req->messages()) {
        REQUIRE_EQ(m.groupid(), 1337);
    }
}

TEST_CASE("batch case: status objects start unfinished, from an app id or from a request", "[util][batch][cases]")
{
    auto byId = batchExecStatusFactory(1337);
    REQUIRE_EQ(byId->appid(), 1337);
    REQUIRE(!byId->finished());
    auto req = batchExecFactory("foo", "bar", 1);
    auto byReq = batchExecStatusFactory(req);
    REQUIRE_EQ(byReq->appid(), req->appid());
    REQUIRE(!byReq->finished());
}

TEST_CASE("batch case: migrated messages do not count as finished", "[util][batch][cases]")
{
    auto req = batchExecFactory("foo", "bar", 4);
    auto status = batchExecStatusFactory(req);
    REQUIRE_EQ(getNumFinishedMessagesInBatch(status), 0);
    *status->add_messageresults() = req->messages(0);
    *status->add_messageresults() = req->messages(1);
    REQUIRE_EQ(getNumFinishedMessagesInBatch(status), 2);
    auto* migrated = status->add_messageresults();
    *migrated = req->messages(2);
    migrated->set_returnvalue(MIGRATED_FUNCTION_RETURN_VALUE);
    REQUIRE_EQ(getNumFinishedMessagesInBatch(status), 2);
}

// ---- tokens ----------------------------------------------------------------
TEST_CASE("token case: tokens come out in order and a released one is reused last", "[util][tokens][cases]")
{
    TokenPool pool(5);
    REQUIRE_EQ(pool.getToken(), 0);
    REQUIRE_EQ(pool.getToken(), 1);
    REQUIRE_EQ(pool.getToken(), 2);
    pool.releaseToken(1);
    REQUIRE_EQ(pool.getToken(), 3);
    REQUIRE_EQ(pool.getToken(), 4);
    REQUIRE_EQ(pool.getToken(), 1);
}

TEST_CASE("token case: a pool shared by threads hands each token out once", "[util][tokens][cases]")
{
    TokenPool pool(3);
    std::mutex mx;
    std::vector<int> acquired;
    auto take = [&] {
        int t = pool.getToken();
        std::lock_guard<std::mutex> lk(mx);
        acquired.push_back(t);
    };
    for (int i = 0; i < 3; i++) {
        std::thread t(take);
        t.join();
    }
    REQUIRE(acquired == (std::vector<int>{ 0, 1, 2 }));
    pool.releaseToken(1);
    std::thread t(take);
    t.join();
    REQUIRE(acquired == (std::vector<int>{ 0, 1, 2, 1 }));
}

TEST_CASE("token case: size, free and taken counts", "[util][tokens][cases]")
{
    TokenPool p(10);
    REQUIRE(p.size() == 10 && p.free() == 10 && p.taken() == 0);
    int a = p.getToken(), b = p.getToken();
    REQUIRE(p.size() == 10 && p.free() == 8 && p.taken() == 2);
    p.releaseToken(a);
    REQUIRE(p.free() == 9 && p.taken() == 1);
    p.releaseToken(b);
    REQUIRE(p.free() == 10 && p.taken() == 0);
}

TEST_CASE("token case: an empty pool answers -1", "[util][tokens][cases]")
{
    TokenPool p(0);
    REQUIRE_EQ(p.getToken(), -1);
    REQUIRE_EQ(p.getToken(), -1);
    REQUIRE_EQ(p.getToken(), -1);
}

// ---- environment / config --------------------------------------------------
TEST_CASE("environment case: an unset variable yields the default", "[util][environment][cases]")
{
    REQUIRE(getenv("JUNK_VAR") == nullptr);
    REQUIRE_EQ(getEnvVar("JUNK_VAR", "blah"), std::string("blah"));
}

TEST_CASE("environment case: setting a variable returns what it held before", "[util][environment][cases]")
{
    unsetEnvVar("MY_VAR");
    REQUIRE(getenv("MY_VAR") == nullptr);
    REQUIRE_EQ(getEnvVar("MY_VAR", "alpha"), std::string("alpha"));
    REQUIRE_EQ(setEnvVar("MY_VAR", "beta"), std::string(""));
    REQUIRE_EQ(setEnvVar("MY_VAR", "gamma"), std::string("beta"));
    REQUIRE_EQ(getEnvVar("MY_VAR", "alpha"), std::string("gamma"));
    unsetEnvVar("MY_VAR");
}

TEST_CASE("environment case: the usable core count can be overridden", "[util][environment][cases]")
{
    auto& conf = getSystemConfig();
    conf.reset();
    unsigned int cores = getUsableCores();
    REQUIRE(cores > 0);
    conf.overrideCpuCount = 1234;
    REQUIRE_EQ(getUsableCores(), 1234u);
    conf.reset();
    REQUIRE_EQ(getUsableCores(), cores);
}

TEST_CASE("config case: defaults", "[util][config][cases]")
{
    // (a scratch object: the process-wide one may carry the harness's settings)
    std::vector<std::pair<std::string, std::string>> saved;
    for (const char* k : { "LOG_LEVEL", "LOG_FILE", "STATE_MODE", "REDIS_PORT", "BATCH_SCHEDULER_MODE", "GLOBAL_MESSAGE_TIMEOUT",
                           "BOUND_TIMEOUT", "DEFAULT_MPI_WORLD_SIZE", "DIRTY_TRACKING_MODE", "PLANNER_PORT" }) {
        saved.emplace_back(k, getEnvVar(k, "\x01unset"));
        unsetEnvVar(k);
    }
    SystemConfig conf;
    conf.reset();
    REQUIRE_EQ(conf.logLevel, std::string("info"));
    REQUIRE_EQ(conf.logFile, std::string("off"));
    REQUIRE_EQ(conf.stateMode, std::string("inmemory"));
    REQUIRE_EQ(conf.redisPort, std::string("6379"));
    REQUIRE_EQ(conf.batchSchedulerMode, std::string("bin-pack"));
    REQUIRE_EQ(conf.globalMessageTimeout, 60000);
    REQUIRE_EQ(conf.boundTimeout, 30000);
    REQUIRE_EQ(conf.defaultMpiWorldSize, 5);
    REQUIRE_EQ(conf.dirtyTrackingMode, std::string("segfault"));
    REQUIRE_EQ(conf.plannerPort, 8080);
    for (auto& [k, v] : saved) {
        if (v != "\x01unset") {
            setEnvVar(k, v);
        }
    }
}

TEST_CASE("config case: every knob follows its environment variable", "[util][config][cases]")
{
    std::vector<std::pair<std::string, std::string>> wanted = {
        { "LOG_LEVEL", "debug" },          { "LOG_FILE", "on" },
        { "STATE_MODE", "foobar" },        { "REDIS_STATE_HOST", "not-localhost" },
        { "REDIS_QUEUE_HOST", "other-host" }, { "REDIS_PORT", "1234" },
        { "OVERRIDE_CPU_COUNT", "4" },     { "BATCH_SCHEDULER_MODE", "foo-bar" },
        { "GLOBAL_MESSAGE_TIMEOUT", "9876" }, { "BOUND_TIMEOUT", "6666" },
        { "FUNCTION_SERVER_THREADS", "111" }, { "STATE_SERVER_THREADS", "222" },
        { "SNAPSHOT_SERVER_THREADS", "333" }, { "POINT_TO_POINT_SERVER_THREADS", "444" },
        { "DEFAULT_MPI_WORLD_SIZE", "2468" }, { "DIRTY_TRACKING_MODE", "dummy-track" },
        { "PLANNER_HOST", "dummy-planner" }, { "PLANNER_PORT", "9876" },
    };
    std::vector<std::pair<std::string, std::string>> saved;
    for (auto& [k, v] : wanted) {
        saved.emplace_back(k, getEnvVar(k, "\x01unset"));
        setEnvVar(k, v);
    }
    SystemConfig conf;
    conf.reset();
    REQUIRE_EQ(conf.logLevel, std::string("debug"));
    REQUIRE_EQ(conf.logFile, std::string("on"));
    REQUIRE_EQ(conf.stateMode, std::string("foobar"));
    REQUIRE_EQ(conf.redisStateHost, std::string("not-localhost"));
    REQUIRE_EQ(conf.redisQueueHost, std::string("other-host"));
    REQUIRE_EQ(conf.redisPort, std::string("1234"));
    REQUIRE_EQ(conf.overrideCpuCount, 4);
    REQUIRE_EQ(conf.batchSchedulerMode, std::string("foo-bar"));
    REQUIRE_EQ(conf.globalMessageTimeout, 9876);
    REQUIRE_EQ(conf.boundTimeout, 6666);
    REQUIRE_EQ(conf.functionServerThreads, 111);
    REQUIRE_EQ(conf.stateServerThreads, 222);
    REQUIRE_EQ(conf.snapshotServerThreads, 333);
    REQUIRE_EQ(conf.pointToPointServerThreads, 444);
    REQUIRE_EQ(conf.defaultMpiWorldSize, 2468);
    REQUIRE_EQ(conf.dirtyTrackingMode, std::string("dummy-track"));
    REQUIRE_EQ(conf.plannerHost, std::string("dummy-planner"));
    REQUIRE_EQ(conf.plannerPort, 9876);
    for (auto& [k, v] : saved) {
        if (v == "\x01unset") {
            unsetEnvVar(k);
        } else {
            setEnvVar(k, v);
        }
    }
    getSystemConfig().reset();
}

// ---- latch / barrier / flag ------------------------------------------------
TEST_CASE("latch case: releases when the last of its count arrives, refuses latecomers", "[util][sync][cases]")
{
    auto l = Latch::create(3);
    std::thread t1([l] { l->wait(); });
    std::thread t2([l] { l->wait(); });
    l->wait();
    t1.join();
    t2.join();
    REQUIRE_THROWS(l->wait());
}

TEST_CASE("latch case: a latch nobody completes times out", "[util][sync][cases]")
{
    auto l = Latch::create(2, 200);
    REQUIRE_THROWS(l->wait());
}

TEST_CASE("barrier case: reusable, with a completion hook run once per cycle", "[util][sync][cases]")
{
    std::atomic<int> completions{ 0 };
    auto b = Barrier::create(3, [&] { completions++; });
    std::atomic<int> phase{ 0 };
    std::atomic<bool> early{ false };
    auto member = [&] {
        for (int r = 0; r < 4; r++) {
            phase++;
            b->wait();
            if (phase.load() < (r + 1) * 3) {
                early = true;
            }
            b->wait();
        }
    };
    std::thread t1(member), t2(member);
    member();
    t1.join();
    t2.join();
    REQUIRE(!early.load());
    REQUIRE_EQ(completions.load(), 8);
}

TEST_CASE("flag case: waiters of one flag are released together, the other flag's keep waiting", "[util][sync][cases]")
{
    const int n = 10;
    auto flagA = std::make_shared<FlagWaiter>();
    auto flagB = std::make_shared<FlagWaiter>();
    auto a1 = Latch::create(n + 1), b1 = Latch::create(n + 1), a2 = Latch::create(n + 1), b2 = Latch::create(n + 1);
    std::vector<int> resultsA(n, 0), resultsB(n, 0), unset(n, 0), set;
    std::vector<std::thread> threads;
    for (int i = 0; i < n; i++) {
        set.push_back(i);
        threads.emplace_back([&, i] {
            a1->wait();
            flagA->waitOnFlag();
            resultsA[i] = i;
            a2->wait();
        });
        threads.emplace_back([&, i] {
            b1->wait();
            flagB->waitOnFlag();
            resultsB[i] = i;
            b2->wait();
        });
    }
    a1->wait();
    b1->wait();
    REQUIRE(resultsA == unset);
    REQUIRE(resultsB == unset);
    flagA->setFlag(true);
    a2->wait();
    REQUIRE(resultsA == set);
    REQUIRE(resultsB == unset);
    flagB->setFlag(true);
    b2->wait();
    REQUIRE(resultsB == set);
    for (auto& t : threads) {
        t.join();
    }
}

// ---- random / gids -----------------------------------------------------------
TEST_CASE("random case: strings of the asked length that differ from each other", "[util][random][cases]")
{
    std::string a = randomString(100), b = randomString(100);
    REQUIRE_EQ(a.size(), 100u);
    REQUIRE_EQ(b.size(), 100u);
    REQUIRE(a != b);
}

TEST_CASE("random case: a random member of a set; nothing from an empty one", "[util][random][cases]")
{
    std::unordered_set<std::string> s;
    REQUIRE(randomStringFromSet(s).empty());
    s = { "foo", "bar", "baz", "qux" };
    std::unordered_set<std::string> seen;
    for (int i = 0; i < 1000; i++) {
        seen.insert(randomStringFromSet(s));
    }
    REQUIRE_EQ(seen.size(), 4u);
}

TEST_CASE("gid case: ids generated by many threads never collide", "[util][gids][cases]")
{
    const int nThreads = 10, perThread = 1000;
    std::vector<std::vector<unsigned int>> ids(nThreads);
    std::vector<std::thread> threads;
    for (int t = 0; t < nThreads; t++) {
        threads.emplace_back([&, t] {
            for (int i = 0; i < perThread; i++) {
                ids[t].push_back(generateGid());
            }
        });
    }
    for (auto& t : threads) {
        t.join();
    }
    std::set<unsigned int> all;
    for (auto& v : ids) {
        all.insert(v.begin(), v.end());
    }
    REQUIRE_EQ(all.size(), (size_t)nThreads * perThread);
    REQUIRE(all.count(0) == 0);
}

// ---- delta settings ----------------------------------------------------------
TEST_CASE("delta case: settings strings, every form", "[util][delta][cases]")
{
    DeltaSettings empty("");
    REQUIRE(!empty.usePages && !empty.xorWithOld && !empty.useZstd);
    for (const char* def : { "pages=64;", "pages=64" }) {
        DeltaSettings s(def);
        REQUIRE(s.usePages && s.pageSize == 64 && !s.xorWithOld && !s.useZstd);
    }
    for (const char* def : { "xor;", "xor" }) {
        DeltaSettings s(def);
        REQUIRE(!s.usePages && s.xorWithOld && !s.useZstd);
    }
    DeltaSettings z("zstd=-3;");
    REQUIRE(!z.usePages && !z.xorWithOld && z.useZstd && z.zstdLevel == -3);
    DeltaSettings z2("zstd=7");
    REQUIRE(z2.useZstd && z2.zstdLevel == 7);
    DeltaSettings all("pages=4096;xor;zstd=1");
    REQUIRE(all.usePages && all.pageSize == 4096 && all.xorWithOld && all.useZstd && all.zstdLevel == 1);
    // what it prints parses back to the same settings
    DeltaSettings again(all.toString());
    REQUIRE(again.usePages && again.pageSize == 4096 && again.xorWithOld && again.useZstd && again.zstdLevel == 1);
    REQUIRE_THROWS(DeltaSettings("bogus=1;"));
}

// ---- hwloc ------------------------------------------------------------------
TEST_CASE("hwloc case: a thread is pinned to a free CPU, released with its handle", "[util][hwloc][cases]")
{
    setTestMode(true);
    pthread_t self = pthread_self();
    cpu_set_t before;
    pthread_getaffinity_np(self, sizeof(before), &before);
    {
        auto cpu = pinThreadToFreeCpu(self);
        REQUIRE(cpu != nullptr);
        REQUIRE(cpu->get() != nullptr);
        cpu_set_t now;
        REQUIRE_EQ(pthread_getaffinity_np(self, sizeof(now), &now), 0);
        REQUIRE_EQ(CPU_COUNT(&now), 1);
    }
    pthread_setaffinity_np(self, sizeof(before), &before);
}

TEST_CASE("hwloc case: more pins than CPUs fail, except in test mode", "[util][hwloc][cases]")
{
    pthread_t self = pthread_self();
    cpu_set_t before;
    pthread_getaffinity_np(self, sizeof(before), &before);
    int nCpus = (int)getUsableCores();
    {
        std::atomic<bool> stop{ false };
        std::vector<std::thread> threads;
        // joins on every way out, a failed check included
        std::shared_ptr<void> joiner(nullptr, [&](void*) {
            stop = true;
            for (auto& t : threads) {
                if (t.joinable()) {
                    t.join();
                }
            }
            setTestMode(true);
        });
        std::vector<std::unique_ptr<FaabricCpuSet>> held;
        setTestMode(false);
        for (int i = 0; i < nCpus; i++) {
            threads.emplace_back([&] {
                while (!stop.load()) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(2));
                }
            });
            held.push_back(pinThreadToFreeCpu(threads.back().native_handle()));
        }
        // every CPU is taken now
        REQUIRE_THROWS(pinThreadToFreeCpu(self));
        setTestMode(true);
        auto extra = pinThreadToFreeCpu(self); // allowed: shares a CPU
        REQUIRE(extra != nullptr);
        cpu_set_t now;
        REQUIRE_EQ(pthread_getaffinity_np(self, sizeof(now), &now), 0);
        REQUIRE(CPU_ISSET(0, &now));
    }
    pthread_setaffinity_np(self, sizeof(before), &before);
    setTestMode(true);
}

// ---- json ---------------------------------------------------------------------
namespace {
faabric::Message jsonMessage()
{
    faabric::Message msg;
    msg.set_type(faabric::Message_MessageType_FLUSH);
    msg.set_user("user 1");
    msg.set_function("great function");
    msg.set_executedhost("blah.host.blah");
    msg.set_finishtimestamp(123456543);
    msg.set_pythonuser("py user");
    msg.set_pythonfunction("py func");
    msg.set_pythonentry("py entry");
    msg.set_ispython(true);
    msg.set_ismpi(true);
    msg.set_mpiworldid(1234);
    msg.set_mpirank(5678);
    msg.set_mpiworldsize(33);
    msg.set_cmdline("some cmdline");
    msg.set_recordexecgraph(true);
    (*msg.mutable_execgraphdetails())["foo"] = "bar";
    (*msg.mutable_intexecgraphdetails())["foo"] = 0;
    msg.set_inputdata("foo bar");
    setMessageId(msg);
    return msg;
}

void checkRoundTrip(const faabric::Message& msg)
{
    faabric::Message back;
    jsonToMessage(messageToJson(msg), &back);
    REQUIRE_EQ(back.id(), msg.id());
    REQUIRE_EQ(back.appid(), msg.appid());
    REQUIRE(back.type() == msg.type());
    REQUIRE_EQ(back.user(), msg.user());
    REQUIRE_EQ(back.function(), msg.function());
    REQUIRE_EQ(back.executedhost(), msg.executedhost());
    REQUIRE_EQ(back.starttimestamp(), msg.starttimestamp());
    REQUIRE_EQ(back.finishtimestamp(), msg.finishtimestamp());
    REQUIRE_EQ(back.pythonuser(), msg.pythonuser());
    REQUIRE_EQ(back.pythonfunction(), msg.pythonfunction());
    REQUIRE_EQ(back.pythonentry(), msg.pythonentry());
    REQUIRE_EQ(back.ispython(), msg.ispython());
    REQUIRE_EQ(back.ismpi(), msg.ismpi());
    REQUIRE_EQ(back.mpiworldid(), msg.mpiworldid());
    REQUIRE_EQ(back.mpirank(), msg.mpirank());
    REQUIRE_EQ(back.mpiworldsize(), msg.mpiworldsize());
    REQUIRE_EQ(back.cmdline(), msg.cmdline());
    REQUIRE_EQ(back.recordexecgraph(), msg.recordexecgraph());
    REQUIRE_EQ(back.inputdata(), msg.inputdata());
    REQUIRE_EQ(back.resultkey(), msg.resultkey());
    REQUIRE_EQ(back.statuskey(), msg.statuskey());
    REQUIRE(back.execgraphdetails().at("foo") == "bar");
    REQUIRE_EQ(back.intexecgraphdetails().at("foo"), 0);
}
}

TEST_CASE("json case: a message survives the round trip, odd characters included", "[util][json][cases]")
{
    faabric::Message msg = jsonMessage();
    REQUIRE(msg.id() > 0);
    REQUIRE(msg.starttimestamp() > 0);
    checkRoundTrip(msg);
    msg.set_inputdata("[0], %$ 2233 9");
    checkRoundTrip(msg);
    msg.set_inputdata("quote \" backslash \\ newline \n tab \t done");
    checkRoundTrip(msg);
}

TEST_CASE("json case: binary input data survives the round trip", "[util][json][cases]")
{
    faabric::Message msg = jsonMessage();
    std::vector<uint8_t> bytes = { 0, 0, 1, 1, 0, 2, 2, 3, 3, 4, 4 };
    msg.set_inputdata(std::string((const char*)bytes.data(), bytes.size()));
    checkRoundTrip(msg);
}

TEST_CASE("json case: the keys other tools read are present, the type is a number", "[util][json][cases]")
{
    std::string json = messageToJson(jsonMessage());
    for (const char* key : { "input_data", "python", "py_user", "py_func", "mpi", "mpi_world_size", "record_exec_graph", "start_ts",
                             "finish_ts" }) {
        REQUIRE(json.find("\"" + std::string(key) + "\":") != std::string::npos);
    }
    REQUIRE(json.find("\"type\":3,") != std::string::npos);
}

// ---- files / network ------------------------------------------------------------
TEST_CASE("files case: bytes written to a file read back the same", "[util][files][cases]")
{
    std::string path = "/tmp/faabric_b200_case_" + std::to_string(generateGid()) + ".txt";
    std::vector<uint8_t> bytes = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 0, 0, 255 };
    writeBytesToFile(path, bytes);
    REQUIRE(readFileToBytes(path) == bytes);
    REQUIRE_EQ(readFileToString(path).size(), bytes.size());
    ::unlink(path.c_str());
    REQUIRE_THROWS(readFileToBytes(path));
}

TEST_CASE("network case: this host has a primary address that is not loopback", "[util][network][cases]")
{
    std::string ip = getPrimaryIPForThisHost("");
    REQUIRE(!ip.empty());
    REQUIRE(ip != "127.0.0.1");
    REQUIRE_EQ(std::count(ip.begin(), ip.end(), '.'), 3);
    REQUIRE_EQ(getIPFromHostname("localhost"), std::string("127.0.0.1"));
}
