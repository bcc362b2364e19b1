// Control-plane endpoints.  The reference wraps NNG sockets (push/pull,
// req/rep, inproc pairs: include/faabric/transport/MessageEndpoint.h:14-280);
// here they are plain TCP streams of framed messages plus an in-process
// registry, because on one box most peers live in the same process: a send to
// a server registered in this process skips the socket entirely.
// No payload of the data plane travels here (that is NVLink P2P).
#pragma once

#include <faabric/transport/Message.h>
#include <faabric/transport/tcp/Socket.h>
#include <faabric/util/exception.h>
#include <faabric/util/queue.h>

#include <memory>
#include <mutex>
#include <string>

#define DEFAULT_SEND_TIMEOUT_MS 60000
#define DEFAULT_RECV_TIMEOUT_MS 60000
#define DEFAULT_SOCKET_TIMEOUT_MS 60000

namespace faabric::transport {

class MessageTimeoutException final : public faabric::util::FaabricException
{
  public:
    explicit MessageTimeoutException(std::string message)
      : FaabricException(std::move(message))
    {}
};

class MessageEndpointServer;

// Blocking helpers over a connected stream socket
void sendFrame(int fd,
               uint8_t code,
               const uint8_t* data,
               size_t size,
               int sequenceNum);

// Returns a Message with response code TERM if the peer closed, TIMEOUT on
// timeout (timeoutMs <= 0: wait forever)
Message recvFrame(int fd, int timeoutMs);

// Base: remembers where it points and lazily connects
class SendMessageEndpoint
{
  public:
    SendMessageEndpoint(const std::string& hostIn, int portIn, int timeoutMsIn);

    virtual ~SendMessageEndpoint();

    std::string getAddress() const { return host + ":" + std::to_string(port); }

    const std::string& getHost() const { return host; }

    int getPort() const { return port; }

  protected:
    std::string host;
    int port;
    int timeoutMs;
    std::unique_ptr<tcp::SendSocket> sock;
    std::mutex sockMx;

    // nullptr when the destination is not served from this process
    MessageEndpointServer* findLocalServer(bool sync);

    int connectedFd();

    void dropConnection();
};

// Fire-and-forget (PUSH-like)
class AsyncSendMessageEndpoint final : public SendMessageEndpoint
{
  public:
    AsyncSendMessageEndpoint(const std::string& hostIn,
                             int portIn,
                             int timeoutMs = DEFAULT_SEND_TIMEOUT_MS);

    void send(uint8_t header,
              const uint8_t* data,
              size_t dataSize,
              int sequenceNum = NO_SEQUENCE_NUM);
};

// Request / response (REQ-like)
class SyncSendMessageEndpoint final : public SendMessageEndpoint
{
  public:
    SyncSendMessageEndpoint(const std::string& hostIn,
                            int portIn,
                            int timeoutMs = DEFAULT_SEND_TIMEOUT_MS);

    void sendRaw(const uint8_t* data, size_t dataSize);

    Message sendAwaitResponse(uint8_t header,
                              const uint8_t* data,
                              size_t dataSize);
};

// In-process mailboxes addressed by label: the local leg of point-to-point
// messaging (reference: inproc:// push/pull and pair endpoints)
class InprocMailbox
{
  public:
    void send(Message&& msg) { queue.enqueue(std::move(msg)); }

    // Throws MessageTimeoutException
    Message recv(int timeoutMs);

    long size() { return queue.size(); }

  private:
    faabric::util::Queue<Message> queue;
};

std::shared_ptr<InprocMailbox> getInprocMailbox(const std::string& label);

void clearInprocMailbox(const std::string& label);

void clearAllInprocMailboxes();

class AsyncInternalSendMessageEndpoint final
{
  public:
    explicit AsyncInternalSendMessageEndpoint(
      const std::string& inprocLabel,
      int timeoutMs = DEFAULT_SEND_TIMEOUT_MS);

    void send(uint8_t header,
              const uint8_t* data,
              size_t dataSize,
              int sequenceNum = NO_SEQUENCE_NUM);

  private:
    std::shared_ptr<InprocMailbox> mailbox;
};

class AsyncInternalRecvMessageEndpoint final
{
  public:
    explicit AsyncInternalRecvMessageEndpoint(
      const std::string& inprocLabel,
      int timeoutMsIn = DEFAULT_RECV_TIMEOUT_MS);

    Message recv();

  private:
    std::shared_ptr<InprocMailbox> mailbox;
    int timeoutMs;
};

// Direct pair: same thing under the names the reference uses
using AsyncDirectSendEndpoint = AsyncInternalSendMessageEndpoint;
using AsyncDirectRecvEndpoint = AsyncInternalRecvMessageEndpoint;

}
