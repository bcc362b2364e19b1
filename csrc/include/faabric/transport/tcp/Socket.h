// Raw TCP sockets: the cross-process MPI fallback data plane for host buffers
// and the carrier of the control RPC (reference: src/transport/tcp/*).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#define SocketListenBacklog 1024
#define SocketPollTimeoutMs 5000
// 16 MiB send / receive buffers
#define SocketBufferSizeBytes (16 * 1024 * 1024)

namespace faabric::transport::tcp {

void setReuseAddr(int fd);
void setNoDelay(int fd);
void setQuickAck(int fd);
void setBusyPolling(int fd);
void setNonBlocking(int fd);
void setBlocking(int fd);
bool isNonBlocking(int fd);
void setRecvTimeoutMs(int fd, int timeoutMs);
void setSendTimeoutMs(int fd, int timeoutMs);
void setRecvBufferSize(int fd, size_t bufferSize);
void setSendBufferSize(int fd, size_t bufferSize);

class Socket
{
  public:
    Socket();
    explicit Socket(int connFd);
    Socket(const Socket&) = delete;
    Socket& operator=(const Socket&) = delete;
    Socket(Socket&& other) noexcept;
    Socket& operator=(Socket&& other) noexcept;
    ~Socket();

    int get() const { return fd; }
    void close();

  private:
    int fd = -1;
};

class SendSocket
{
  public:
    SendSocket(const std::string& hostIn, int portIn);

    // Retries while the peer is still coming up (30 x 200 ms by default)
    void dial(int retries = 30, int sleepMs = 200);

    void sendOne(const uint8_t* buffer, size_t bufferSize);

    int getFd() const { return sock.get(); }

  private:
    std::string host;
    int port;
    Socket sock;
};

class RecvSocket
{
  public:
    explicit RecvSocket(int portIn, const std::string& hostIn = "0.0.0.0");
    ~RecvSocket();

    void listen();

    // Returns the connection fd; throws on timeout
    int accept(int timeoutMs = SocketPollTimeoutMs);

    void recvOne(int conn, uint8_t* buffer, size_t bufferSize);

    int getPort() const { return port; }

  private:
    std::string host;
    int port;
    Socket sock;
    std::vector<int> openConnections;
};

}
