// Bare receiving endpoints under the reference's names
// (reference: tests/test/transport/test_message_endpoint_client.cpp - send/recv
// one and many messages, await a response, timeouts, many senders to one
// receiver; API at include/faabric/transport/MessageEndpoint.h:158-253)
#include "fixtures.h"

#include <faabric/transport/MessageEndpoint.h>

#include <thread>

using namespace tests;
using namespace faabric::transport;

namespace {

constexpr int PORT_A = 23481;

std::string textOf(const Message& m)
{
    auto d = m.udata();
    return std::string((const char*)d.data(), d.size());
}
}

TEST_CASE("recv endpoints: async messages arrive in order with header and sequence number", "[transport][endpoints]")
{
    AsyncRecvMessageEndpoint recv(PORT_A, 2000);
    AsyncSendMessageEndpoint send(LOCALHOST, PORT_A);
    for (int i = 0; i < 5; i++) {
        std::string body = "message-" + std::to_string(i);
        send.send((uint8_t)(10 + i), BYTES_CONST(body.data()), body.size(), 100 + i);
    }
    for (int i = 0; i < 5; i++) {
        Message m = recv.recv();
        REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
        REQUIRE_EQ((int)m.getMessageCode(), 10 + i);
        REQUIRE_EQ(m.getSequenceNum(), 100 + i);
        REQUIRE_EQ(textOf(m), "message-" + std::to_string(i));
    }
}

TEST_CASE("recv endpoints: an empty port times out, a stopped one terminates", "[transport][endpoints]")
{
    AsyncRecvMessageEndpoint recv(PORT_A + 1, 50);
    Message m = recv.recv();
    REQUIRE(m.getResponseCode() == MessageResponseCode::TIMEOUT);
    recv.stop();
    REQUIRE(recv.recv().getResponseCode() == MessageResponseCode::TERM);
    // an empty body is a valid message
    AsyncRecvMessageEndpoint other(PORT_A + 2, 2000);
    AsyncSendMessageEndpoint send(LOCALHOST, PORT_A + 2);
    send.send(7, nullptr, 0);
    Message e = other.recv();
    REQUIRE(e.getResponseCode() == MessageResponseCode::SUCCESS);
    REQUIRE_EQ((int)e.getMessageCode(), 7);
    REQUIRE_EQ(e.size(), 0u);
}

TEST_CASE("recv endpoints: a sync receiver answers on the sender's connection", "[transport][endpoints]")
{
    SyncRecvMessageEndpoint recv(PORT_A + 3, 5000);
    std::thread server([&] {
        for (int i = 0; i < 3; i++) {
            Message m = recv.recv();
            std::string reply = "echo:" + textOf(m);
            recv.sendResponse(m.getMessageCode(), BYTES_CONST(reply.data()), reply.size());
        }
    });
    SyncSendMessageEndpoint send(LOCALHOST, PORT_A + 3, 5000);
    for (int i = 0; i < 3; i++) {
        std::string body = "req" + std::to_string(i);
        Message res = send.sendAwaitResponse((uint8_t)(20 + i), BYTES_CONST(body.data()), body.size());
        REQUIRE_EQ((int)res.getMessageCode(), 20 + i);
        REQUIRE_EQ(textOf(res), "echo:" + body);
    }
    server.join();
    // nobody answers any more: the client's wait runs out
    SyncSendMessageEndpoint impatient(LOCALHOST, PORT_A + 3, 100);
    REQUIRE_THROWS(impatient.sendAwaitResponse(1, BYTES_CONST("x"), 1));
}

TEST_CASE("recv endpoints: many senders into one receiver", "[transport][endpoints]")
{
    AsyncRecvMessageEndpoint recv(PORT_A + 4, 5000);
    const int nSenders = 6, perSender = 50;
    std::vector<std::thread> senders;
    for (int s = 0; s < nSenders; s++) {
        senders.emplace_back([s] {
            AsyncSendMessageEndpoint send(LOCALHOST, PORT_A + 4);
            for (int i = 0; i < perSender; i++) {
                int v[2] = { s, i };
                send.send(3, BYTES_CONST(v), sizeof(v));
            }
        });
    }
    // per-sender order is kept, all of them arrive
    std::vector<int> nextOf(nSenders, 0);
    for (int k = 0; k < nSenders * perSender; k++) {
        Message m = recv.recv();
        REQUIRE(m.getResponseCode() == MessageResponseCode::SUCCESS);
        const int* v = (const int*)m.udata().data();
        REQUIRE_EQ(v[1], nextOf[v[0]]);
        nextOf[v[0]]++;
    }
    for (auto& t : senders) {
        t.join();
    }
    for (int s = 0; s < nSenders; s++) {
        REQUIRE_EQ(nextOf[s], perSender);
    }
}

TEST_CASE("recv endpoints: a sync fan endpoint shares requests between attached workers", "[transport][endpoints]")
{
    SyncFanMessageEndpoint fan(PORT_A + 5, 5000);
    const int nWorkers = 3, nClients = 4, perClient = 20;
    std::atomic<int> handled[nWorkers] = {};
    std::atomic<int> terminated{ 0 };
    std::vector<std::thread> workers;
    for (int w = 0; w < nWorkers; w++) {
        workers.emplace_back([&] {
            MessageContext ctx = fan.attachFanOut();
            while (true) {
                Message m = fan.recv(ctx);
                if (m.getResponseCode() == MessageResponseCode::TERM) {
                    terminated++;
                    return;
                }
                if (m.getResponseCode() != MessageResponseCode::SUCCESS) {
                    continue;
                }
                int v = *(const int*)m.udata().data() * 2;
                handled[ctx.getWorkerId()]++;
                fan.sendResponse(ctx, m.getMessageCode(), BYTES_CONST(&v), sizeof(v));
            }
        });
    }
    std::atomic<int> bad{ 0 };
    std::vector<std::thread> clients;
    for (int c = 0; c < nClients; c++) {
        clients.emplace_back([&, c] {
            SyncSendMessageEndpoint send(LOCALHOST, PORT_A + 5, 5000);
            for (int i = 0; i < perClient; i++) {
                int v = c * 1000 + i;
                Message res = send.sendAwaitResponse(9, BYTES_CONST(&v), sizeof(v));
                if (res.size() != sizeof(int) || *(const int*)res.udata().data() != 2 * v) {
                    bad++;
                }
            }
        });
    }
    for (auto& t : clients) {
        t.join();
    }
    REQUIRE_EQ(bad.load(), 0);
    int total = 0;
    for (int w = 0; w < nWorkers; w++) {
        total += handled[w].load();
    }
    REQUIRE_EQ(total, nClients * perClient);
    fan.stop();
    for (auto& t : workers) {
        t.join();
    }
    REQUIRE_EQ(terminated.load(), nWorkers);
    // an unattached context is refused, an async fan does not respond
    REQUIRE_THROWS(fan.recv(MessageContext()));
    AsyncFanMessageEndpoint afan(PORT_A + 6, 100);
    MessageContext ctx = afan.attachFanOut();
    REQUIRE(afan.recv(ctx).getResponseCode() == MessageResponseCode::TIMEOUT);
    REQUIRE_THROWS(afan.sendResponse(ctx, 1, nullptr, 0));
}

TEST_CASE("recv endpoints: in-process labels, async and direct pairs", "[transport][endpoints]")
{
    AsyncRecvMessageEndpoint recv("endpoint-test-label", 500);
    AsyncInternalSendMessageEndpoint send("endpoint-test-label");
    std::string body = "in-process";
    send.send(4, BYTES_CONST(body.data()), body.size(), 77);
    Message m = recv.recv();
    REQUIRE_EQ(textOf(m), body);
    REQUIRE_EQ(m.getSequenceNum(), 77);
    REQUIRE(recv.recv().getResponseCode() == MessageResponseCode::TIMEOUT);
    // the direct pair is the same mailbox mechanism
    AsyncDirectRecvEndpoint drecv("endpoint-test-direct", 500);
    AsyncDirectSendEndpoint dsend("endpoint-test-direct");
    dsend.send(5, BYTES_CONST(body.data()), 2);
    REQUIRE_EQ(textOf(drecv.recv()), std::string("in"));
    clearInprocMailbox("endpoint-test-label");
    clearInprocMailbox("endpoint-test-direct");
}
