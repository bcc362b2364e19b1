// One case per case of the reference's endpoint and server tests, over real
// TCP (127.0.0.2 is loopback but not "this host", so the in-process fast path
// is bypassed)
// (reference: tests/test/transport/test_message.cpp,
// test_message_endpoint_client.cpp:18-330, test_message_server.cpp:150-362)
#include "harness.h"

#include <faabric/proto/faabric.pb.h>
#include <faabric/transport/Message.h>
#include <faabric/transport/MessageEndpoint.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/util/latch.h>

#include <atomic>
#include <thread>

using namespace faabric::transport;

namespace {
const char* TCP_HOST = "127.0.0.2";
constexpr int PORT = 23611; // endpoint cases: +0..+19
constexpr int SRV_ASYNC = 23651;
constexpr int SRV_SYNC = 23652;

std::string textOf(const Message& m)
{
    auto d = m.udata();
    return std::string((const char*)d.data(), d.size());
}

class CountingServer final : public MessageEndpointServer
{
  public:
    CountingServer()
      : MessageEndpointServer(SRV_ASYNC, SRV_SYNC, "test-dummy", 2)
    {}

    std::atomic<int> messageCount{ 0 };

  protected:
    void doAsyncRecv(Message&) override { messageCount++; }

    std::string doSyncRecv(Message&) override
    {
        messageCount++;
        return faabric::EmptyResponse().SerializeAsString();
    }
};

// echoes the body; an int body is also a number of milliseconds to sleep
// first when `sleeps` is set; `latch` makes every handler wait for a second one
class EchoingServer final : public MessageEndpointServer
{
  public:
    explicit EchoingServer(int nThreads = 2)
      : MessageEndpointServer(SRV_ASYNC, SRV_SYNC, "test-echo", nThreads)
    {}

    bool sleeps = false;
    std::shared_ptr<faabric::util::Latch> latch;

  protected:
    void doAsyncRecv(Message&) override { throw std::runtime_error("not expecting async messages"); }

    std::string doSyncRecv(Message& message) override
    {
        if (sleeps && message.size() == sizeof(int)) {
            std::this_thread::sleep_for(std::chrono::milliseconds(*(const int*)message.udata().data()));
        }
        if (latch) {
            latch->wait();
        }
        faabric::StatePart response;
        response.set_data(std::string((const char*)message.udata().data(), message.size()));
        return response.SerializeAsString();
    }
};
}

TEST_CASE("transport case: moving a message moves its payload, not a copy", "[transport][cases]")
{
    std::vector<uint8_t> payload = { 1, 2, 3, 4 };
    const uint8_t* where = payload.data();
    Message a(7, 42, std::move(payload));
    REQUIRE(a.udata().data() == where);
    Message b(std::move(a));
    REQUIRE(b.udata().data() == where);
    REQUIRE_EQ((int)b.getMessageCode(), 7);
    REQUIRE_EQ(b.getSequenceNum(), 42);
    REQUIRE_EQ(b.size(), 4u);
    Message c;
    c = std::move(b);
    REQUIRE(c.udata().data() == where);
    REQUIRE(c.dataCopy() == (std::vector<uint8_t>{ 1, 2, 3, 4 }));
}

TEST_CASE("transport case: send / recv one message", "[transport][cases]")
{
    AsyncRecvMessageEndpoint recv(PORT, 2000);
    AsyncSendMessageEndpoint send(TCP_HOST, PORT);
    std::string body = "Hello world!";
    send.send(3, BYTES_CONST(body.data()), body.size());
    Message m = recv.recv();
    REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
    REQUIRE_EQ((int)m.getMessageCode(), 3);
    REQUIRE_EQ(textOf(m), body);
}

TEST_CASE("transport case: a send issued before the receiver exists arrives once it does", "[transport][cases]")
{
    std::string body = "Hello world!";
    std::thread sender([&] {
        AsyncSendMessageEndpoint send(TCP_HOST, PORT + 1);
        send.send(0, BYTES_CONST(body.data()), body.size()); // dials until somebody listens
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(300));
    AsyncRecvMessageEndpoint recv(PORT + 1, 5000);
    Message m = recv.recv();
    sender.join();
    REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
    REQUIRE_EQ(textOf(m), body);
}

TEST_CASE("transport case: await a response", "[transport][cases]")
{
    std::string question = "Hello ", answer = "world!";
    std::thread server([&] {
        SyncRecvMessageEndpoint recv(PORT + 2, 5000);
        Message m = recv.recv();
        if (textOf(m) == question) {
            recv.sendResponse(0, BYTES_CONST(answer.data()), answer.size());
        }
    });
    SyncSendMessageEndpoint send(TCP_HOST, PORT + 2, 5000);
    Message res = send.sendAwaitResponse(0, BYTES_CONST(question.data()), question.size());
    server.join();
    REQUIRE_EQ(textOf(res), answer);
}

TEST_CASE("transport case: send / recv many messages", "[transport][cases]")
{
    const int n = 10000;
    AsyncRecvMessageEndpoint recv(PORT + 3, 5000);
    std::thread sender([&] {
        AsyncSendMessageEndpoint send(TCP_HOST, PORT + 3);
        std::string body = "Hello world!";
        for (int i = 0; i < n; i++) {
            send.send((uint8_t)(i % 200), BYTES_CONST(body.data()), body.size(), i);
        }
    });
    for (int i = 0; i < n; i++) {
        Message m = recv.recv();
        REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
        REQUIRE_EQ(m.getSequenceNum(), i);
        REQUIRE_EQ(m.size(), 12u);
    }
    sender.join();
}

TEST_CASE("transport case: send / recv many messages from many clients", "[transport][cases]")
{
    const int nClients = 10, perClient = 1000;
    AsyncRecvMessageEndpoint recv(PORT + 4, 5000);
    std::vector<std::thread> senders;
    for (int c = 0; c < nClients; c++) {
        senders.emplace_back([&] {
            AsyncSendMessageEndpoint send(TCP_HOST, PORT + 4);
            std::string body = "Hello world!";
            for (int i = 0; i < perClient; i++) {
                send.send(0, BYTES_CONST(body.data()), body.size());
            }
        });
    }
    for (int i = 0; i < nClients * perClient; i++) {
        Message m = recv.recv();
        REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
        REQUIRE_EQ(textOf(m), std::string("Hello world!"));
    }
    for (auto& t : senders) {
        t.join();
    }
}

TEST_CASE("transport case: zero and negative timeouts are refused by every endpoint", "[transport][cases]")
{
    {
        // sanity: valid ones are fine
        AsyncSendMessageEndpoint s(TCP_HOST, PORT + 5, 100);
        AsyncRecvMessageEndpoint r(PORT + 5, 100);
        SyncSendMessageEndpoint sb(TCP_HOST, PORT + 6, 100);
        SyncRecvMessageEndpoint rb(PORT + 6, 100);
    }
    for (int bad : { 0, -1 }) {
        REQUIRE_THROWS(AsyncRecvMessageEndpoint(PORT + 5, bad));
        REQUIRE_THROWS(SyncRecvMessageEndpoint(PORT + 6, bad));
        REQUIRE_THROWS(AsyncSendMessageEndpoint(TCP_HOST, PORT + 5, bad));
        REQUIRE_THROWS(SyncSendMessageEndpoint(TCP_HOST, PORT + 6, bad));
        REQUIRE_THROWS(AsyncFanMessageEndpoint(PORT + 7, bad));
    }
}

TEST_CASE("transport case: direct messaging between two threads of one process", "[transport][cases]")
{
    const std::string label = "direct-case";
    std::string expected = "Direct hello";
    std::atomic<bool> ok{ false };
    std::thread receiver([&] {
        AsyncDirectRecvEndpoint recv(label, 5000);
        ok = textOf(recv.recv()) == expected;
    });
    AsyncDirectSendEndpoint send(label);
    send.send(0, BYTES_CONST(expected.data()), expected.size());
    receiver.join();
    REQUIRE(ok.load());
    clearInprocMailbox(label);
}

TEST_CASE("transport case: direct messaging stress, many pairs at once", "[transport][cases]")
{
    const int nPairs = 20, perPair = 500;
    std::atomic<int> good{ 0 };
    std::vector<std::thread> threads;
    for (int p = 0; p < nPairs; p++) {
        std::string label = "direct-stress-" + std::to_string(p);
        threads.emplace_back([label, &good] {
            AsyncDirectRecvEndpoint recv(label, 10000);
            for (int i = 0; i < perPair; i++) {
                Message m = recv.recv();
                if (m.getSequenceNum() == i && m.size() == sizeof(int) && *(const int*)m.udata().data() == i) {
                    good++;
                }
            }
        });
        threads.emplace_back([label] {
            AsyncDirectSendEndpoint send(label);
            for (int i = 0; i < perPair; i++) {
                send.send(1, BYTES_CONST(&i), sizeof(i), i);
            }
        });
    }
    for (auto& t : threads) {
        t.join();
    }
    REQUIRE_EQ(good.load(), nPairs * perPair);
    clearAllInprocMailboxes();
}

TEST_CASE("transport case: one message to a server", "[transport][cases]")
{
    CountingServer server;
    server.start();
    REQUIRE_EQ(server.messageCount.load(), 0);
    MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC);
    std::string body = "body";
    server.setRequestLatch();
    cli.asyncSend(0, BYTES_CONST(body.data()), body.size());
    server.awaitRequestLatch();
    REQUIRE_EQ(server.messageCount.load(), 1);
    server.stop();
}

TEST_CASE("transport case: a server's response reaches the client", "[transport][cases]")
{
    EchoingServer server;
    server.start();
    MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC);
    std::string expected = "Response from server";
    faabric::StatePart response;
    cli.syncSend(0, BYTES_CONST(expected.data()), expected.size(), &response);
    REQUIRE_EQ(response.data(), expected);
    server.stop();
}

TEST_CASE("transport case: many clients talk to one server", "[transport][cases]")
{
    CountingServer server;
    server.start();
    const int nClients = 10, perClient = 1000;
    std::vector<std::thread> clients;
    for (int c = 0; c < nClients; c++) {
        clients.emplace_back([&] {
            MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC);
            std::string body = "Message from threaded client";
            for (int i = 0; i < perClient; i++) {
                cli.asyncSend(0, BYTES_CONST(body.data()), body.size());
            }
        });
    }
    for (auto& t : clients) {
        t.join();
    }
    for (int i = 0; i < 1000 && server.messageCount.load() < nClients * perClient; i++) {
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    REQUIRE_EQ(server.messageCount.load(), nClients * perClient);
    server.stop();
}

TEST_CASE("transport case: a client's timeout: long enough for one handler, too short for another", "[transport][cases]")
{
    EchoingServer server;
    server.sleeps = true;
    server.start();
    {
        // the handler takes 100 ms, the client gives it 1000
        MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC, 1000);
        int delay = 100;
        faabric::StatePart response;
        cli.syncSend(0, BYTES_CONST(&delay), sizeof(delay), &response);
        REQUIRE_EQ(response.data().size(), sizeof(int));
    }
    {
        // the handler takes 1000 ms, the client gives it 100
        MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC, 100);
        int delay = 1000;
        faabric::StatePart response;
        bool timedOut = false;
        try {
            cli.syncSend(0, BYTES_CONST(&delay), sizeof(delay), &response);
        } catch (const MessageTimeoutException&) {
            timedOut = true;
        }
        REQUIRE(timedOut);
    }
    std::this_thread::sleep_for(std::chrono::milliseconds(1000)); // let the slow handler finish
    server.stop();
}

TEST_CASE("transport case: blocking requests are served by different threads of the server", "[transport][cases]")
{
    EchoingServer server(2);
    server.latch = faabric::util::Latch::create(2);
    server.start();
    std::atomic<int> good{ 0 };
    auto ask = [&](const std::string& text) {
        MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC);
        faabric::StatePart response;
        cli.syncSend(0, BYTES_CONST(text.data()), text.size(), &response);
        if (response.data() == text) {
            good++;
        }
    };
    // neither handler returns until both are running
    std::thread a(ask, "Background thread A");
    std::thread b(ask, "Background thread B");
    a.join();
    b.join();
    REQUIRE_EQ(good.load(), 2);
    server.stop();
}

TEST_CASE("transport case: a server keeps listening after its socket timeout passed", "[transport][cases]")
{
    const int timeoutMs = 100;
    CountingServer server;
    server.start(timeoutMs);
    MessageEndpointClient cli(TCP_HOST, SRV_ASYNC, SRV_SYNC);
    std::string body = "body";
    server.setRequestLatch();
    cli.asyncSend(0, BYTES_CONST(body.data()), body.size());
    server.awaitRequestLatch();
    REQUIRE_EQ(server.messageCount.load(), 1);
    std::this_thread::sleep_for(std::chrono::milliseconds(5 * timeoutMs));
    server.setRequestLatch();
    cli.asyncSend(0, BYTES_CONST(body.data()), body.size());
    server.awaitRequestLatch();
    REQUIRE_EQ(server.messageCount.load(), 2);
    server.stop();
}

TEST_CASE("transport case: socket addresses", "[transport][cases]")
{
    faabric::transport::tcp::Address a("127.0.0.2", 8123);
    REQUIRE_EQ(a.port(), 8123);
    REQUIRE_EQ(a.host(), std::string("127.0.0.2"));
    REQUIRE_EQ(((sockaddr_in*)a.get())->sin_family, (sa_family_t)AF_INET);
    faabric::transport::tcp::Address any(9000);
    REQUIRE_EQ(any.port(), 9000);
    REQUIRE_EQ(any.host(), std::string("0.0.0.0"));
    REQUIRE_THROWS(faabric::transport::tcp::Address("not-an-ip", 1));
}

// ---- raw TCP sockets (reference: tests/test/transport/test_tcp_sockets.cpp) ----
#include <faabric/transport/tcp/RecvSocket.h>
#include <faabric/transport/tcp/SendSocket.h>
#include <faabric/transport/tcp/SocketOptions.h>

#include <fcntl.h>

namespace {
const int RAW_PORT = 9977;

// fd numbers are reused quickly: a closed one is recognised by fcntl failing
bool isClosed(int fd)
{
    return ::fcntl(fd, F_GETFD) == -1 && errno == EBADF;
}

void rawSendRecv(const std::vector<int>& msg, bool sendNothing)
{
    namespace tcp = faabric::transport::tcp;
    tcp::RecvSocket dst(RAW_PORT);
    auto done = faabric::util::Latch::create(2);
    std::thread sender([&] {
        tcp::SendSocket src("127.0.0.1", RAW_PORT);
        src.dial();
        if (!sendNothing) {
            src.sendOne(BYTES_CONST(msg.data()), sizeof(int) * msg.size());
        }
        done->wait();
    });
    dst.listen();
    int conn = dst.accept();
    std::vector<int> actual(msg.size());
    if (sendNothing) {
        tcp::setRecvTimeoutMs(conn, 200);
        REQUIRE_THROWS(dst.recvOne(conn, BYTES(actual.data()), sizeof(int) * actual.size()));
    } else {
        tcp::setRecvBufferSize(conn, SocketBufferSizeBytes);
        dst.recvOne(conn, BYTES(actual.data()), sizeof(int) * actual.size());
        REQUIRE(actual == msg);
    }
    done->wait();
    sender.join();
}
}

TEST_CASE("tcp case: an accepted connection is closed with its receiving socket", "[transport][tcp][cases]")
{
    namespace tcp = faabric::transport::tcp;
    int conn = -1;
    {
        tcp::RecvSocket dst(RAW_PORT);
        std::thread sender([&] {
            tcp::SendSocket src("127.0.0.1", RAW_PORT);
            src.dial();
        });
        dst.listen();
        conn = dst.accept();
        REQUIRE(conn >= 0);
        REQUIRE(!isClosed(conn));
        sender.join();
    }
    REQUIRE(isClosed(conn));
}

TEST_CASE("tcp case: every socket option applies to an open connection and throws on a closed one", "[transport][tcp][cases]")
{
    namespace tcp = faabric::transport::tcp;
    auto done = faabric::util::Latch::create(2);
    int conn = -1;
    {
        tcp::RecvSocket dst(RAW_PORT);
        std::thread sender([&] {
            tcp::SendSocket src("127.0.0.1", RAW_PORT);
            src.dial();
            done->wait();
        });
        dst.listen();
        conn = dst.accept();
        tcp::setReuseAddr(conn);
        tcp::setNoDelay(conn);
        tcp::setQuickAck(conn);
        tcp::setQuickAck(conn);
        tcp::setBusyPolling(conn);
        tcp::setNonBlocking(conn);
        REQUIRE(tcp::isNonBlocking(conn));
        tcp::setBlocking(conn);
        tcp::setRecvTimeoutMs(conn, 5000);
        tcp::setSendTimeoutMs(conn, 5000);
        tcp::setRecvBufferSize(conn, SocketBufferSizeBytes);
        tcp::setSendBufferSize(conn, SocketBufferSizeBytes);
        REQUIRE(!tcp::isNonBlocking(conn));
        done->wait();
        sender.join();
    }
    REQUIRE(isClosed(conn));
    REQUIRE_THROWS(tcp::setReuseAddr(conn));
    REQUIRE_THROWS(tcp::setNoDelay(conn));
    REQUIRE_THROWS(tcp::setQuickAck(conn));
    REQUIRE_THROWS(tcp::setBusyPolling(conn));
    REQUIRE_THROWS(tcp::setNonBlocking(conn));
    REQUIRE_THROWS(tcp::setBlocking(conn));
    REQUIRE_THROWS(tcp::setRecvTimeoutMs(conn, 5000));
    REQUIRE_THROWS(tcp::setSendTimeoutMs(conn, 5000));
    REQUIRE_THROWS(tcp::setRecvBufferSize(conn, SocketBufferSizeBytes));
    REQUIRE_THROWS(tcp::setSendBufferSize(conn, SocketBufferSizeBytes));
}

TEST_CASE("tcp case: one small message over raw sockets", "[transport][tcp][cases]")
{
    rawSendRecv(std::vector<int>(3, 2), false);
}

TEST_CASE("tcp case: one large message over raw sockets", "[transport][tcp][cases]")
{
    rawSendRecv(std::vector<int>(300, 200), false);
}

TEST_CASE("tcp case: receiving times out when nothing is sent", "[transport][tcp][cases]")
{
    rawSendRecv(std::vector<int>(3, 2), true);
}
