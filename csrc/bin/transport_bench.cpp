// Control-plane microbenchmarks: RPC round trips (in-process fast path vs TCP
// loopback), async message rate, and local point-to-point messaging.
// Prints one JSON line.
#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/transport/PointToPointBroker.h>
#include <faabric/transport/PointToPointServer.h>
#include <faabric/util/config.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>

#include <atomic>
#include <chrono>
#include <thread>

using namespace faabric::transport;

class EchoServer final : public MessageEndpointServer
{
  public:
    EchoServer()
      : MessageEndpointServer(9811, 9812, "bench-echo", 4)
    {}

    std::atomic<long> asyncCount{ 0 };

  protected:
    void doAsyncRecv(Message& message) override { asyncCount++; }

    std::string doSyncRecv(Message& message) override
    {
        return std::string((const char*)message.udata().data(), message.size());
    }
};

static double rttUs(const std::string& host, size_t bytes, int iters)
{
    MessageEndpointClient cli(host, 9811, 9812);
    std::vector<uint8_t> payload(bytes, 7);
    for (int i = 0; i < iters / 10 + 10; i++) {
        cli.syncSendRaw(1, payload.data(), payload.size());
    }
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < iters; i++) {
        cli.syncSendRaw(1, payload.data(), payload.size());
    }
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / iters;
}

static double asyncRate(EchoServer& server, const std::string& host, int n)
{
    MessageEndpointClient cli(host, 9811, 9812);
    uint8_t b[32] = { 0 };
    long before = server.asyncCount.load();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) {
        cli.asyncSend(2, b, sizeof(b));
    }
    while (server.asyncCount.load() < before + n) {
        std::this_thread::yield();
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return n / s;
}

int main()
{
    setenv("LOG_LEVEL", "warn", 0);
    faabric::util::getSystemConfig().reset();
    faabric::util::initLogging();
    EchoServer server;
    server.start();
    const std::string tcpHost = "127.0.0.2"; // loopback, but not "this host": forces TCP

    printf("{\"bench\": \"transport\"");
    for (size_t bytes : { (size_t)8, (size_t)65536, (size_t)1 << 20 }) {
        int iters = bytes > 100000 ? 500 : 5000;
        printf(", \"rtt_inproc_us_%zu\": %.2f", bytes, rttUs(LOCALHOST, bytes, iters));
        printf(", \"rtt_tcp_us_%zu\": %.2f", bytes, rttUs(tcpHost, bytes, iters));
    }
    printf(", \"async_inproc_msgs_per_s\": %.0f", asyncRate(server, LOCALHOST, 200000));
    printf(", \"async_tcp_msgs_per_s\": %.0f", asyncRate(server, tcpHost, 200000));
    server.stop();

    // Point-to-point messaging between two group members on this host
    auto& broker = getPointToPointBroker();
    PointToPointServer ptpServer;
    ptpServer.start();
    faabric::batch_scheduler::SchedulingDecision d(1, 2);
    std::string thisHost = faabric::util::getSystemConfig().endpointHost;
    d.addMessage(thisHost, 10, 0, 0);
    d.addMessage(thisHost, 11, 1, 1);
    broker.setAndSendMappingsFromSchedulingDecision(d);
    const int n = 20000;
    std::thread peer([&] {
        for (int i = 0; i < n; i++) {
            auto m = broker.recvMessage(2, 0, 1);
            broker.sendMessage(2, 1, 0, m.data(), m.size());
        }
        broker.resetThreadLocalCache();
    });
    uint8_t token[8] = { 0 };
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; i++) {
        broker.sendMessage(2, 0, 1, token, sizeof(token));
        broker.recvMessage(2, 1, 0);
    }
    double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / n;
    peer.join();
    printf(", \"ptp_local_pingpong_us\": %.2f}\n", us);
    broker.clear();
    ptpServer.stop();
    return 0;
}
