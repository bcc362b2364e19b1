// Server skeleton: one async (PULL-like) and one sync (REP-like) listener, N
// worker threads each (reference: src/transport/MessageEndpointServer.cpp:
// 18-230).  All connections of a listener are multiplexed by one epoll I/O
// thread that frames messages and hands them to the workers.  Clients that live
// in the same process bypass the sockets through the server registry.
#pragma once

#include <faabric/transport/Message.h>
#include <faabric/transport/MessageEndpoint.h>
#include <faabric/util/latch.h>
#include <faabric/util/queue.h>

#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace faabric::transport {

// Dedicated thread count when none is given
#define DEFAULT_MESSAGE_SERVER_THREADS 4

class MessageEndpointServer;

class MessageEndpointServerHandler
{
  public:
    MessageEndpointServerHandler(MessageEndpointServer* serverIn,
                                 bool asyncIn,
                                 const std::string& inprocLabelIn,
                                 int nThreadsIn);

    ~MessageEndpointServerHandler();

    void start(int timeoutMs);

    void join();

    // In-process delivery: enqueue for a worker (async)
    void deliverLocal(Message&& msg);

    int getPort() const { return port; }

  private:
    struct Impl;
    std::unique_ptr<Impl> impl;
    MessageEndpointServer* server;
    bool async;
    std::string inprocLabel;
    int nThreads;
    int port = 0;
};

class MessageEndpointServer
{
  public:
    MessageEndpointServer(int asyncPortIn,
                          int syncPortIn,
                          const std::string& inprocLabelIn,
                          int nThreadsIn);

    virtual ~MessageEndpointServer();

    virtual void start(int timeoutMs = DEFAULT_SOCKET_TIMEOUT_MS);

    virtual void stop();

    virtual void onWorkerStop();

    // Test hook: the next request waits on this latch after being handled
    void setRequestLatch();

    void awaitRequestLatch();

    int getNThreads() const { return nThreads; }

    bool isStarted() const { return started.load(); }

    // ---- used by handlers and by the in-process fast path ----
    virtual void doAsyncRecv(transport::Message& message) = 0;

    // Returns the serialised response
    virtual std::string doSyncRecv(transport::Message& message) = 0;

    void handleAsync(Message& msg);

    std::string handleSync(Message& msg);

    MessageEndpointServerHandler* getAsyncHandler() { return &asyncHandler; }

    static MessageEndpointServer* findLocal(int port, bool sync);

  protected:
    int asyncPort;
    int syncPort;
    std::string inprocLabel;
    int nThreads;

  private:
    friend class MessageEndpointServerHandler;

    MessageEndpointServerHandler asyncHandler;
    MessageEndpointServerHandler syncHandler;

    std::atomic<bool> started{ false };
    std::shared_ptr<faabric::util::Latch> requestLatch;
    std::mutex latchMx;

    void afterRequest();
};

// "ip" / "localhost" / this host's address all count as local
bool isLocalAddress(const std::string& host);

}
