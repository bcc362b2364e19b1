// Stand-alone receiving endpoints: a bound port (or an in-process label) whose
// framed messages are handed out by recv().  Same wire format as the servers,
// so the Async/Sync send endpoints talk to them unchanged
// (reference API: include/faabric/transport/MessageEndpoint.h:158-253).
#include <faabric/transport/MessageEndpoint.h>
#include <faabric/transport/tcp/Socket.h>
#include <faabric/util/logging.h>
#include <faabric/util/queue.h>

#include <poll.h>
#include <sys/eventfd.h>
#include <sys/socket.h>
#include <unistd.h>

namespace faabric::transport {

// One listener thread per bound port: accepts connections, reads complete
// frames and queues them with the connection they arrived on
class PortListener
{
  public:
    struct Item
    {
        Message msg;
        int fd = -1;
    };

    explicit PortListener(int portIn)
      : sock(portIn)
    {
        sock.listen();
        wakeFd = ::eventfd(0, EFD_NONBLOCK);
        thread = std::thread([this] { run(); });
    }

    ~PortListener() { stop(); }

    void stop()
    {
        bool expected = false;
        if (!stopping.compare_exchange_strong(expected, true)) {
            return;
        }
        uint64_t one = 1;
        (void)!::write(wakeFd, &one, sizeof(one));
        if (thread.joinable()) {
            thread.join();
        }
        ::close(wakeFd);
        // wake every blocked receiver, now and later
        for (int i = 0; i < 64; i++) {
            queue.enqueue(Item{ Message(MessageResponseCode::TERM), -1 });
        }
    }

    bool stopped() const { return stopping.load(); }

    // TIMEOUT / TERM are reported through the message's response code
    Item next(int timeoutMs)
    {
        if (stopping.load()) {
            return Item{ Message(MessageResponseCode::TERM), -1 };
        }
        try {
            return queue.dequeue(timeoutMs <= 0 ? 3600 * 1000 : timeoutMs);
        } catch (const faabric::util::QueueTimeoutException&) {
            return Item{ Message(MessageResponseCode::TIMEOUT), -1 };
        }
    }

    void respond(int fd, uint8_t header, const uint8_t* data, size_t size, int seq)
    {
        if (fd < 0) {
            throw std::runtime_error("No connection to respond on");
        }
        std::lock_guard<std::mutex> lk(sendMx);
        sendFrame(fd, header, data, size, seq);
    }

  private:
    tcp::RecvSocket sock;
    int wakeFd = -1;
    std::thread thread;
    std::atomic<bool> stopping{ false };
    faabric::util::Queue<Item> queue;
    std::mutex sendMx;

    void run()
    {
        std::vector<int> conns;
        while (!stopping.load()) {
            std::vector<pollfd> fds;
            fds.push_back({ wakeFd, POLLIN, 0 });
            fds.push_back({ sock.getFd(), POLLIN, 0 });
            for (int c : conns) {
                fds.push_back({ c, POLLIN, 0 });
            }
            int pr = ::poll(fds.data(), fds.size(), -1);
            if (pr < 0) {
                if (errno == EINTR) {
                    continue;
                }
                break;
            }
            if (fds[0].revents != 0) {
                break;
            }
            for (size_t i = 2; i < fds.size(); i++) {
                if (fds[i].revents == 0) {
                    continue;
                }
                int fd = fds[i].fd;
                Message m = recvFrame(fd, 5000);
                if (m.getResponseCode() != MessageResponseCode::SUCCESS) {
                    // closed (or a frame that never completed): stop watching
                    conns.erase(std::find(conns.begin(), conns.end(), fd));
                    continue;
                }
                queue.enqueue(Item{ std::move(m), fd });
            }
            if (fds[1].revents != 0) {
                try {
                    conns.push_back(sock.accept(100));
                } catch (const std::exception&) {
                    // the dialler went away again
                }
            }
        }
        // RecvSocket closes the connections it accepted
    }
};

// ---------------------------------------------------------------------------
static int validTimeout(int timeoutMs)
{
    if (timeoutMs <= 0) {
        SPDLOG_ERROR("Setting invalid timeout of {}", timeoutMs);
        throw std::runtime_error("Setting invalid timeout");
    }
    return timeoutMs;
}

RecvMessageEndpoint::RecvMessageEndpoint(int portIn, int timeoutMsIn)
  : port(portIn)
  , timeoutMs(validTimeout(timeoutMsIn))
  , listener(std::make_shared<PortListener>(portIn))
{}

RecvMessageEndpoint::RecvMessageEndpoint(const std::string& inprocLabel, int timeoutMsIn)
  : timeoutMs(validTimeout(timeoutMsIn))
  , mailbox(getInprocMailbox(inprocLabel))
{}

RecvMessageEndpoint::~RecvMessageEndpoint()
{
    stop();
}

void RecvMessageEndpoint::stop()
{
    if (listener != nullptr) {
        listener->stop();
    }
}

Message RecvMessageEndpoint::doRecv(MessageContext& ctx)
{
    if (mailbox != nullptr) {
        try {
            return mailbox->recv(timeoutMs);
        } catch (const MessageTimeoutException&) {
            return Message(MessageResponseCode::TIMEOUT);
        }
    }
    PortListener::Item it = listener->next(timeoutMs);
    ctx.replyFd = it.fd;
    ctx.replySeq = it.msg.getSequenceNum();
    return std::move(it.msg);
}

Message RecvMessageEndpoint::recv()
{
    return doRecv(last);
}

void RecvMessageEndpoint::reply(const MessageContext& ctx, uint8_t header, const uint8_t* data, size_t dataSize)
{
    if (listener == nullptr) {
        throw std::runtime_error("In-process endpoints have no connection to respond on");
    }
    listener->respond(ctx.replyFd, header, data, dataSize, NO_SEQUENCE_NUM);
}

AsyncRecvMessageEndpoint::AsyncRecvMessageEndpoint(int portIn, int timeoutMs)
  : RecvMessageEndpoint(portIn, timeoutMs)
{}

AsyncRecvMessageEndpoint::AsyncRecvMessageEndpoint(const std::string& inprocLabel, int timeoutMs)
  : RecvMessageEndpoint(inprocLabel, timeoutMs)
{}

SyncRecvMessageEndpoint::SyncRecvMessageEndpoint(int portIn, int timeoutMs)
  : RecvMessageEndpoint(portIn, timeoutMs)
{}

void SyncRecvMessageEndpoint::sendResponse(uint8_t header, const uint8_t* data, size_t dataSize)
{
    reply(last, header, data, dataSize);
}

// ---------------------------------------------------------------------------
FanMessageEndpoint::FanMessageEndpoint(int portIn, int timeoutMsIn, bool isAsyncIn)
  : port(portIn)
  , timeoutMs(validTimeout(timeoutMsIn))
  , isAsync(isAsyncIn)
  , listener(std::make_shared<PortListener>(portIn))
{}

FanMessageEndpoint::~FanMessageEndpoint()
{
    stop();
}

MessageContext FanMessageEndpoint::attachFanOut()
{
    MessageContext ctx;
    ctx.workerId = nWorkers.fetch_add(1);
    return ctx;
}

Message FanMessageEndpoint::recv(const MessageContext& ctx)
{
    if (!ctx.isValid()) {
        throw std::runtime_error("Receiving on a fan endpoint without attaching first");
    }
    PortListener::Item it = listener->next(timeoutMs);
    ctx.replyFd = it.fd;
    ctx.replySeq = it.msg.getSequenceNum();
    return std::move(it.msg);
}

void FanMessageEndpoint::sendResponse(const MessageContext& ctx, uint8_t header, const uint8_t* data, size_t dataSize)
{
    if (isAsync) {
        throw std::runtime_error("Async fan endpoints do not respond");
    }
    listener->respond(ctx.replyFd, header, data, dataSize, NO_SEQUENCE_NUM);
}

void FanMessageEndpoint::stop()
{
    listener->stop();
}

}
