#include <algorithm>
#include <faabric/transport/MessageEndpoint.h>
#include <faabric/transport/MessageEndpointClient.h>
#include <faabric/transport/MessageEndpointServer.h>
#include <faabric/transport/common.h>
#include <faabric/util/config.h>
#include <faabric/util/fault.h>
#include <faabric/util/logging.h>
#include <faabric/util/network.h>
#include <faabric/util/string_tools.h>
#include <faabric/util/testing.h>

#include <arpa/inet.h>
#include <cerrno>
#include <cstring>
#include <map>
#include <netinet/in.h>
#include <poll.h>
#include <shared_mutex>
#include <thread>
#include <sys/epoll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>
#include <unordered_map>

namespace faabric::transport {

// ---------------------------------------------------------------------------
// Addresses
// ---------------------------------------------------------------------------
// Virtual hosts: one worker process exposes each of its GPUs to the planner as
// a separate "host" (gpu0..gpuN-1); the alias table routes them to the
// process that really serves them.
static std::shared_mutex aliasMx;
static std::unordered_map<std::string, std::string> hostAliases;

void registerHostAlias(const std::string& alias, const std::string& realAddress)
{
    std::unique_lock<std::shared_mutex> lk(aliasMx);
    hostAliases[alias] = realAddress;
}

void clearHostAliases()
{
    std::unique_lock<std::shared_mutex> lk(aliasMx);
    hostAliases.clear();
}

std::string resolveHostAlias(const std::string& host)
{
    std::shared_lock<std::shared_mutex> lk(aliasMx);
    auto it = hostAliases.find(host);
    return it == hostAliases.end() ? host : it->second;
}

bool isHostAlias(const std::string& host)
{
    std::shared_lock<std::shared_mutex> lk(aliasMx);
    return hostAliases.find(host) != hostAliases.end();
}

bool sameWorker(const std::string& hostA, const std::string& hostB)
{
    if (hostA.empty() || hostB.empty()) {
        return false;
    }
    return resolveHostAlias(hostA) == resolveHostAlias(hostB);
}

HostAddress parseHostAddress(const std::string& hostIn)
{
    HostAddress a;
    std::string host = resolveHostAlias(hostIn);
    size_t colon = host.rfind(':');
    if (colon != std::string::npos &&
        faabric::util::stringIsInt(host.substr(colon + 1))) {
        a.ip = host.substr(0, colon);
        a.portOffset = std::stoi(host.substr(colon + 1));
    } else {
        a.ip = host;
    }
    return a;
}

std::string makeHostAddress(const std::string& ip, int portOffset)
{
    if (portOffset == 0) {
        return ip;
    }
    return ip + ":" + std::to_string(portOffset);
}

std::string getThisHostAddress()
{
    auto& conf = faabric::util::getSystemConfig();
    return makeHostAddress(conf.endpointHost, conf.portOffset);
}

bool isLocalAddress(const std::string& host)
{
    if (host == "localhost" || host == LOCALHOST || host == "0.0.0.0") {
        return true;
    }
    return host == faabric::util::getSystemConfig().endpointHost;
}

// ---------------------------------------------------------------------------
// Framing
// ---------------------------------------------------------------------------
static void writeAll(int fd, const uint8_t* p, size_t n)
{
    while (n > 0) {
        ssize_t w = ::send(fd, p, n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) {
                continue;
            }
            if (errno == EAGAIN || errno == EWOULDBLOCK) {
                pollfd pf{ fd, POLLOUT, 0 };
                ::poll(&pf, 1, 1000);
                continue;
            }
            throw std::runtime_error(std::string("Socket send failed: ") + strerror(errno));
        }
        p += w;
        n -= (size_t)w;
    }
}

void sendFrame(int fd,
               uint8_t code,
               const uint8_t* data,
               size_t size,
               int sequenceNum)
{
    uint8_t header[HEADER_MSG_SIZE];
    Message::writeHeader(header, code, size, sequenceNum);
    if (size <= 4096) {
        // One syscall for small messages
        uint8_t buf[HEADER_MSG_SIZE + 4096];
        memcpy(buf, header, HEADER_MSG_SIZE);
        if (size > 0) {
            memcpy(buf + HEADER_MSG_SIZE, data, size);
        }
        writeAll(fd, buf, HEADER_MSG_SIZE + size);
    } else {
        writeAll(fd, header, HEADER_MSG_SIZE);
        writeAll(fd, data, size);
    }
}

// 0 = ok, 1 = closed, 2 = timeout
static int readAll(int fd, uint8_t* p, size_t n, int timeoutMs)
{
    while (n > 0) {
        pollfd pf{ fd, POLLIN, 0 };
        int pr = ::poll(&pf, 1, timeoutMs <= 0 ? -1 : timeoutMs);
        if (pr == 0) {
            return 2;
        }
        if (pr < 0) {
            if (errno == EINTR) {
                continue;
            }
            return 1;
        }
        ssize_t r = ::recv(fd, p, n, 0);
        if (r == 0) {
            return 1;
        }
        if (r < 0) {
            if (errno == EINTR || errno == EAGAIN) {
                continue;
            }
            return 1;
        }
        p += r;
        n -= (size_t)r;
    }
    return 0;
}

Message recvFrame(int fd, int timeoutMs)
{
    uint8_t header[HEADER_MSG_SIZE];
    int rc = readAll(fd, header, HEADER_MSG_SIZE, timeoutMs);
    if (rc == 1) {
        return Message(MessageResponseCode::TERM);
    }
    if (rc == 2) {
        return Message(MessageResponseCode::TIMEOUT);
    }
    uint8_t code;
    uint64_t size;
    int32_t seq;
    Message::readHeader(header, code, size, seq);
    std::vector<uint8_t> payload(size);
    if (size > 0) {
        rc = readAll(fd, payload.data(), size, timeoutMs);
        if (rc == 1) {
            return Message(MessageResponseCode::TERM);
        }
        if (rc == 2) {
            return Message(MessageResponseCode::TIMEOUT);
        }
    }
    return Message(code, seq, std::move(payload));
}

// ---------------------------------------------------------------------------
// Local server registry
// ---------------------------------------------------------------------------
static std::shared_mutex registryMx;
static std::unordered_map<int, MessageEndpointServer*> asyncServers;
static std::unordered_map<int, MessageEndpointServer*> syncServers;

static bool inprocRpcEnabled()
{
    static bool enabled = []() {
        const char* v = getenv("FAABRIC_INPROC_RPC");
        return v == nullptr || std::string(v) != "0";
    }();
    return enabled;
}

MessageEndpointServer* MessageEndpointServer::findLocal(int port, bool sync)
{
    std::shared_lock<std::shared_mutex> lk(registryMx);
    auto& m = sync ? syncServers : asyncServers;
    auto it = m.find(port);
    return it == m.end() ? nullptr : it->second;
}

// ---------------------------------------------------------------------------
// Send endpoints
// ---------------------------------------------------------------------------
SendMessageEndpoint::SendMessageEndpoint(const std::string& hostIn,
                                         int portIn,
                                         int timeoutMsIn)
  : timeoutMs(timeoutMsIn)
{
    // (reference: MessageEndpoint's constructor refuses them too)
    if (timeoutMsIn <= 0) {
        SPDLOG_ERROR("Setting invalid timeout of {}", timeoutMsIn);
        throw std::runtime_error("Setting invalid timeout");
    }
    HostAddress a = parseHostAddress(hostIn);
    host = a.ip;
    port = portIn + a.portOffset;
}

SendMessageEndpoint::~SendMessageEndpoint() = default;

MessageEndpointServer* SendMessageEndpoint::findLocalServer(bool sync)
{
    if (!inprocRpcEnabled() || !isLocalAddress(host)) {
        return nullptr;
    }
    return MessageEndpointServer::findLocal(port, sync);
}

int SendMessageEndpoint::connectedFd()
{
    if (sock == nullptr) {
        std::string target = host;
        if (target == "0.0.0.0") {
            target = LOCALHOST;
        }
        auto s = std::make_unique<tcp::SendSocket>(target, port);
        // Keep retrying while a peer may still be booting, but never for
        // longer than this endpoint's own timeout
        const int sleepMs = 100;
        int retries = std::clamp(timeoutMs / sleepMs, 1, 60);
        s->dial(retries, sleepMs);
        sock = std::move(s);
    }
    return sock->getFd();
}

void SendMessageEndpoint::dropConnection()
{
    sock.reset();
}

// Returns true when the message must be dropped
static bool injectFault(int port, uint8_t header, const std::string& address)
{
    auto& faults = faabric::util::FaultInjector::get();
    if (!faults.armed()) {
        return false;
    }
    auto rule = faults.match(port, header);
    if (!rule.has_value()) {
        return false;
    }
    switch (rule->action) {
        case faabric::util::FaultAction::DROP:
            SPDLOG_WARN("Fault injection: dropping message {} to {}", (int)header, address);
            return true;
        case faabric::util::FaultAction::DELAY:
            std::this_thread::sleep_for(std::chrono::milliseconds(rule->delayMs));
            return false;
        case faabric::util::FaultAction::ERROR:
            throw std::runtime_error("Fault injection: send of message " + std::to_string((int)header) + " to " + address + " failed");
    }
    return false;
}

AsyncSendMessageEndpoint::AsyncSendMessageEndpoint(const std::string& hostIn,
                                                   int portIn,
                                                   int timeoutMs)
  : SendMessageEndpoint(hostIn, portIn, timeoutMs)
{}

void AsyncSendMessageEndpoint::send(uint8_t header,
                                    const uint8_t* data,
                                    size_t dataSize,
                                    int sequenceNum)
{
    if (injectFault(port, header, getAddress())) {
        return;
    }
    if (MessageEndpointServer* local = findLocalServer(false)) {
        local->getAsyncHandler()->deliverLocal(
          Message(header, sequenceNum, data, dataSize));
        return;
    }
    std::lock_guard<std::mutex> lk(sockMx);
    try {
        sendFrame(connectedFd(), header, data, dataSize, sequenceNum);
    } catch (const std::exception&) {
        // One reconnect attempt (the server may have restarted)
        dropConnection();
        sendFrame(connectedFd(), header, data, dataSize, sequenceNum);
    }
}

SyncSendMessageEndpoint::SyncSendMessageEndpoint(const std::string& hostIn,
                                                 int portIn,
                                                 int timeoutMs)
  : SendMessageEndpoint(hostIn, portIn, timeoutMs)
{}

void SyncSendMessageEndpoint::sendRaw(const uint8_t* data, size_t dataSize)
{
    std::lock_guard<std::mutex> lk(sockMx);
    writeAll(connectedFd(), data, dataSize);
}

Message SyncSendMessageEndpoint::sendAwaitResponse(uint8_t header,
                                                   const uint8_t* data,
                                                   size_t dataSize)
{
    if (injectFault(port, header, getAddress())) {
        // The request is lost: what the caller sees is a timeout
        std::this_thread::sleep_for(std::chrono::milliseconds(std::min(timeoutMs, 200)));
        throw MessageTimeoutException("Timed out waiting for response from " + getAddress() + " (injected drop)");
    }
    if (MessageEndpointServer* local = findLocalServer(true)) {
        // Direct call on the caller's thread: no serialisation hop
        Message req = Message::view(header, NO_SEQUENCE_NUM, data, dataSize);
        return Message(NO_HEADER, NO_SEQUENCE_NUM, local->handleSync(req));
    }
    std::lock_guard<std::mutex> lk(sockMx);
    int fd;
    try {
        fd = connectedFd();
        sendFrame(fd, header, data, dataSize, NO_SEQUENCE_NUM);
    } catch (const std::exception&) {
        dropConnection();
        fd = connectedFd();
        sendFrame(fd, header, data, dataSize, NO_SEQUENCE_NUM);
    }
    Message res = recvFrame(fd, timeoutMs);
    if (res.getResponseCode() == MessageResponseCode::TIMEOUT) {
        dropConnection();
        SPDLOG_ERROR("Timed out waiting for response from {}", getAddress());
        throw MessageTimeoutException("Timed out waiting for response from " + getAddress());
    }
    if (res.getResponseCode() != MessageResponseCode::SUCCESS) {
        dropConnection();
        throw std::runtime_error("Connection closed awaiting response from " + getAddress());
    }
    if (res.getMessageCode() == ERROR_HEADER) {
        throw std::runtime_error("Remote handler on " + getAddress() + " failed: " +
                                 std::string((const char*)res.udata().data(), res.size()));
    }
    return res;
}

// ---------------------------------------------------------------------------
// In-process mailboxes
// ---------------------------------------------------------------------------
static std::mutex mailboxMx;
static std::unordered_map<std::string, std::shared_ptr<InprocMailbox>> mailboxes;

Message InprocMailbox::recv(int timeoutMs)
{
    try {
        return queue.dequeue(timeoutMs <= 0 ? 3600 * 1000 : timeoutMs);
    } catch (const faabric::util::QueueTimeoutException&) {
        throw MessageTimeoutException("Timed out receiving in-process message");
    }
}

std::shared_ptr<InprocMailbox> getInprocMailbox(const std::string& label)
{
    std::lock_guard<std::mutex> lk(mailboxMx);
    auto& slot = mailboxes[label];
    if (slot == nullptr) {
        slot = std::make_shared<InprocMailbox>();
    }
    return slot;
}

void clearInprocMailbox(const std::string& label)
{
    std::lock_guard<std::mutex> lk(mailboxMx);
    mailboxes.erase(label);
}

void clearAllInprocMailboxes()
{
    std::lock_guard<std::mutex> lk(mailboxMx);
    mailboxes.clear();
}

AsyncInternalSendMessageEndpoint::AsyncInternalSendMessageEndpoint(
  const std::string& inprocLabel,
  int timeoutMs)
  : mailbox(getInprocMailbox(inprocLabel))
{}

void AsyncInternalSendMessageEndpoint::send(uint8_t header,
                                            const uint8_t* data,
                                            size_t dataSize,
                                            int sequenceNum)
{
    mailbox->send(Message(header, sequenceNum, data, dataSize));
}

AsyncInternalRecvMessageEndpoint::AsyncInternalRecvMessageEndpoint(
  const std::string& inprocLabel,
  int timeoutMsIn)
  : mailbox(getInprocMailbox(inprocLabel))
  , timeoutMs(timeoutMsIn)
{}

Message AsyncInternalRecvMessageEndpoint::recv()
{
    return mailbox->recv(timeoutMs);
}

// ---------------------------------------------------------------------------
// Client
// ---------------------------------------------------------------------------
MessageEndpointClient::MessageEndpointClient(std::string hostIn,
                                             int asyncPortIn,
                                             int syncPortIn,
                                             int timeoutMs)
  : host(std::move(hostIn))
  , asyncPort(asyncPortIn)
  , syncPort(syncPortIn)
  , asyncEndpoint(host, asyncPort, timeoutMs)
  , syncEndpoint(host, syncPort, timeoutMs)
{}

void MessageEndpointClient::asyncSend(int header,
                                      const uint8_t* buffer,
                                      size_t bufferSize,
                                      int sequenceNum)
{
    asyncEndpoint.send((uint8_t)header, buffer, bufferSize, sequenceNum);
}

Message MessageEndpointClient::syncSendRaw(int header,
                                           const uint8_t* buffer,
                                           size_t bufferSize)
{
    return syncEndpoint.sendAwaitResponse((uint8_t)header, buffer, bufferSize);
}

// ---------------------------------------------------------------------------
// Server handler: epoll I/O thread + worker pool
// ---------------------------------------------------------------------------
struct WorkItem
{
    int fd = -1; // connection to answer on (sync), -1 for in-process
    Message msg;
    bool poison = false;
    std::function<void()> task; // typed in-process request
};

struct Conn
{
    std::vector<uint8_t> buf;
};

struct MessageEndpointServerHandler::Impl
{
    int listenFd = -1;
    int epollFd = -1;
    int stopFd = -1;
    std::thread ioThread;
    std::vector<std::thread> workers;
    faabric::util::Queue<std::shared_ptr<WorkItem>> work;
    std::unordered_map<int, Conn> conns;
    std::atomic<bool> running{ false };
};

MessageEndpointServerHandler::MessageEndpointServerHandler(
  MessageEndpointServer* serverIn,
  bool asyncIn,
  const std::string& inprocLabelIn,
  int nThreadsIn)
  : impl(std::make_unique<Impl>())
  , server(serverIn)
  , async(asyncIn)
  , inprocLabel(inprocLabelIn)
  , nThreads(nThreadsIn)
{}

MessageEndpointServerHandler::~MessageEndpointServerHandler()
{
    join();
}

void MessageEndpointServerHandler::deliverLocal(Message&& msg)
{
    auto item = std::make_shared<WorkItem>();
    item->msg = std::move(msg);
    impl->work.enqueue(std::move(item));
}

void MessageEndpointServerHandler::deliverLocalTask(std::function<void()> task)
{
    auto item = std::make_shared<WorkItem>();
    item->task = std::move(task);
    impl->work.enqueue(std::move(item));
}

void MessageEndpointServerHandler::start(int timeoutMs)
{
    port = async ? server->asyncPort : server->syncPort;
    Impl& im = *impl;

    im.listenFd = ::socket(AF_INET, SOCK_STREAM | SOCK_NONBLOCK, 0);
    if (im.listenFd < 0) {
        throw std::runtime_error("Could not create server socket");
    }
    tcp::setReuseAddr(im.listenFd);
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
    addr.sin_port = htons((uint16_t)port);
    int rc = -1;
    for (int attempt = 0; attempt < 10 && rc != 0; attempt++) {
        rc = ::bind(im.listenFd, (sockaddr*)&addr, sizeof(addr));
        if (rc != 0) {
            std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
    }
    if (rc != 0) {
        ::close(im.listenFd);
        im.listenFd = -1;
        SPDLOG_ERROR("Could not bind {} server to port {}: {}", inprocLabel, port, strerror(errno));
        throw std::runtime_error("Could not bind server socket");
    }
    ::listen(im.listenFd, SocketListenBacklog);

    im.epollFd = ::epoll_create1(0);
    im.stopFd = ::eventfd(0, EFD_NONBLOCK);
    epoll_event ev;
    memset(&ev, 0, sizeof(ev));
    ev.events = EPOLLIN;
    ev.data.fd = im.listenFd;
    ::epoll_ctl(im.epollFd, EPOLL_CTL_ADD, im.listenFd, &ev);
    ev.data.fd = im.stopFd;
    ::epoll_ctl(im.epollFd, EPOLL_CTL_ADD, im.stopFd, &ev);
    im.running.store(true);

    // ---- I/O thread ----
    im.ioThread = std::thread([this] {
        Impl& im = *impl;
        std::vector<epoll_event> events(64);
        std::vector<uint8_t> chunk(256 * 1024);
        while (im.running.load()) {
            int n = ::epoll_wait(im.epollFd, events.data(), (int)events.size(), 500);
            for (int i = 0; i < n; i++) {
                int fd = events[i].data.fd;
                if (fd == im.stopFd) {
                    continue;
                }
                if (fd == im.listenFd) {
                    while (true) {
                        int c = ::accept4(im.listenFd, nullptr, nullptr, SOCK_NONBLOCK);
                        if (c < 0) {
                            break;
                        }
                        try {
                            tcp::setNoDelay(c);
                        } catch (const std::exception&) {
                            // (a peer that is already gone: the read reports it)
                        }
                        epoll_event cev;
                        memset(&cev, 0, sizeof(cev));
                        cev.events = EPOLLIN;
                        cev.data.fd = c;
                        ::epoll_ctl(im.epollFd, EPOLL_CTL_ADD, c, &cev);
                        im.conns[c];
                    }
                    continue;
                }
                // Connection readable: drain, frame, dispatch
                Conn& conn = im.conns[fd];
                bool closed = false;
                while (true) {
                    ssize_t r = ::recv(fd, chunk.data(), chunk.size(), 0);
                    if (r > 0) {
                        conn.buf.insert(conn.buf.end(), chunk.data(), chunk.data() + r);
                        if ((size_t)r < chunk.size()) {
                            break;
                        }
                    } else if (r == 0) {
                        closed = true;
                        break;
                    } else {
                        if (errno == EINTR) {
                            continue;
                        }
                        if (errno != EAGAIN && errno != EWOULDBLOCK) {
                            closed = true;
                        }
                        break;
                    }
                }
                size_t pos = 0;
                while (conn.buf.size() - pos >= HEADER_MSG_SIZE) {
                    uint8_t code;
                    uint64_t size;
                    int32_t seq;
                    Message::readHeader(conn.buf.data() + pos, code, size, seq);
                    if (conn.buf.size() - pos - HEADER_MSG_SIZE < size) {
                        break;
                    }
                    auto item = std::make_shared<WorkItem>();
                    item->fd = async ? -1 : fd;
                    item->msg = Message(code,
                                        seq,
                                        conn.buf.data() + pos + HEADER_MSG_SIZE,
                                        (size_t)size);
                    im.work.enqueue(std::move(item));
                    pos += HEADER_MSG_SIZE + size;
                }
                if (pos > 0) {
                    conn.buf.erase(conn.buf.begin(), conn.buf.begin() + pos);
                }
                if (closed) {
                    ::epoll_ctl(im.epollFd, EPOLL_CTL_DEL, fd, nullptr);
                    ::close(fd);
                    im.conns.erase(fd);
                }
            }
        }
        for (auto& [fd, c] : im.conns) {
            ::close(fd);
        }
        im.conns.clear();
    });

    // ---- workers ----
    for (int t = 0; t < nThreads; t++) {
        im.workers.emplace_back([this] {
            Impl& im = *impl;
            // After serving a request a worker polls briefly for the next one
            // before it blocks: bursts (1024 results of a fan-out, streams of
            // point-to-point messages) otherwise pay a futex wake per message
            bool hot = false;
            while (true) {
                std::shared_ptr<WorkItem> item;
                if (hot) {
                    auto pollStart = std::chrono::steady_clock::now();
                    do {
                        im.work.dequeueIfPresent(&item);
                        if (item != nullptr) {
                            break;
                        }
                        std::this_thread::yield();
                    } while (std::chrono::steady_clock::now() - pollStart < std::chrono::microseconds(30));
                }
                if (item == nullptr) {
                    hot = false;
                    try {
                        item = im.work.dequeue(1000);
                    } catch (const faabric::util::QueueTimeoutException&) {
                        if (!im.running.load()) {
                            break;
                        }
                        continue;
                    }
                }
                hot = true;
                if (item->poison) {
                    break;
                }
                try {
                    if (item->task) {
                        server->runAsyncTask(item->task);
                    } else if (async) {
                        server->handleAsync(item->msg);
                    } else {
                        std::string resp = server->handleSync(item->msg);
                        if (item->fd >= 0) {
                            sendFrame(item->fd,
                                      NO_HEADER,
                                      (const uint8_t*)resp.data(),
                                      resp.size(),
                                      NO_SEQUENCE_NUM);
                        }
                    }
                } catch (const std::exception& e) {
                    SPDLOG_ERROR("Error in {} {} server handler: {}",
                                 inprocLabel,
                                 async ? "async" : "sync",
                                 e.what());
                    if (!async && item->fd >= 0) {
                        // Always answer, or the client hangs until timeout
                        try {
                            std::string what = e.what();
                            sendFrame(item->fd,
                                      ERROR_HEADER,
                                      (const uint8_t*)what.data(),
                                      what.size(),
                                      NO_SEQUENCE_NUM);
                        } catch (...) {
                        }
                    }
                }
            }
            server->onWorkerStop();
        });
    }
}

void MessageEndpointServerHandler::join()
{
    Impl& im = *impl;
    if (!im.running.exchange(false)) {
        return;
    }
    uint64_t one = 1;
    if (::write(im.stopFd, &one, sizeof(one)) < 0) {
        // the epoll timeout ends the loop anyway
    }
    for (size_t i = 0; i < im.workers.size(); i++) {
        auto item = std::make_shared<WorkItem>();
        item->poison = true;
        im.work.enqueue(std::move(item));
    }
    for (auto& w : im.workers) {
        if (w.joinable()) {
            w.join();
        }
    }
    im.workers.clear();
    if (im.ioThread.joinable()) {
        im.ioThread.join();
    }
    if (im.listenFd >= 0) {
        ::close(im.listenFd);
    }
    if (im.epollFd >= 0) {
        ::close(im.epollFd);
    }
    if (im.stopFd >= 0) {
        ::close(im.stopFd);
    }
    im.listenFd = im.epollFd = im.stopFd = -1;
    im.work.reset();
}

// ---------------------------------------------------------------------------
// Server
// ---------------------------------------------------------------------------
MessageEndpointServer::MessageEndpointServer(int asyncPortIn,
                                             int syncPortIn,
                                             const std::string& inprocLabelIn,
                                             int nThreadsIn)
  : asyncPort(asyncPortIn + faabric::util::getSystemConfig().portOffset)
  , syncPort(syncPortIn + faabric::util::getSystemConfig().portOffset)
  , inprocLabel(inprocLabelIn)
  , nThreads(nThreadsIn)
  , asyncHandler(this, true, inprocLabelIn + "-async", nThreadsIn)
  , syncHandler(this, false, inprocLabelIn + "-sync", nThreadsIn)
{}

MessageEndpointServer::~MessageEndpointServer()
{
    stop();
}

void MessageEndpointServer::start(int timeoutMs)
{
    if (started.exchange(true)) {
        return;
    }
    asyncHandler.start(timeoutMs);
    syncHandler.start(timeoutMs);
    std::unique_lock<std::shared_mutex> lk(registryMx);
    asyncServers[asyncPort] = this;
    syncServers[syncPort] = this;
}

void MessageEndpointServer::stop()
{
    if (!started.exchange(false)) {
        return;
    }
    {
        std::unique_lock<std::shared_mutex> lk(registryMx);
        auto a = asyncServers.find(asyncPort);
        if (a != asyncServers.end() && a->second == this) {
            asyncServers.erase(a);
        }
        auto s = syncServers.find(syncPort);
        if (s != syncServers.end() && s->second == this) {
            syncServers.erase(s);
        }
    }
    asyncHandler.join();
    syncHandler.join();
}

void MessageEndpointServer::onWorkerStop() {}

void MessageEndpointServer::handleAsync(Message& msg)
{
    doAsyncRecv(msg);
    afterRequest();
}

void MessageEndpointServer::runAsyncTask(const std::function<void()>& task)
{
    task();
    afterRequest();
}

MessageEndpointServer* MessageEndpointServer::localServerFor(const std::string& host, int basePort, bool sync)
{
    if (!inprocRpcEnabled()) {
        return nullptr;
    }
    HostAddress a = parseHostAddress(host);
    if (!isLocalAddress(a.ip)) {
        return nullptr;
    }
    return findLocal(basePort + a.portOffset, sync);
}

std::string MessageEndpointServer::handleSync(Message& msg)
{
    std::string resp = doSyncRecv(msg);
    afterRequest();
    return resp;
}

void MessageEndpointServer::setRequestLatch()
{
    std::lock_guard<std::mutex> lk(latchMx);
    requestLatch = faabric::util::Latch::create(2);
}

void MessageEndpointServer::awaitRequestLatch()
{
    std::shared_ptr<faabric::util::Latch> l;
    {
        std::lock_guard<std::mutex> lk(latchMx);
        l = requestLatch;
    }
    if (l != nullptr) {
        l->wait();
        std::lock_guard<std::mutex> lk(latchMx);
        requestLatch = nullptr;
    }
}

void MessageEndpointServer::afterRequest()
{
    std::shared_ptr<faabric::util::Latch> l;
    {
        std::lock_guard<std::mutex> lk(latchMx);
        l = requestLatch;
    }
    if (l != nullptr) {
        l->wait();
    }
}

} // namespace faabric::transport
