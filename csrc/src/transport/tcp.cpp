#include <faabric/transport/tcp/Socket.h>
#include <faabric/util/logging.h>

#include <arpa/inet.h>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <stdexcept>
#include <sys/socket.h>
#include <thread>
#include <unistd.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define CPU_RELAX() _mm_pause()
#else
#define CPU_RELAX() std::this_thread::yield()
#endif

namespace faabric::transport::tcp {

// Options throw when they cannot be applied (reference:
// src/transport/tcp/SocketOptions.cpp does for every one of them).  The two
// latency hints may be refused by policy in a sandbox (no CAP_NET_ADMIN, a
// socket family without them): that alone is only logged.
static void setOpt(int fd, int level, int name, int value, const char* what, bool hint = false)
{
    if (::setsockopt(fd, level, name, &value, sizeof(value)) != 0) {
        int err = errno;
        if (hint && (err == EPERM || err == ENOPROTOOPT || err == EOPNOTSUPP)) {
            SPDLOG_DEBUG("setsockopt {} not applied on fd {}: {}", what, fd, strerror(err));
            return;
        }
        SPDLOG_ERROR("setsockopt {} failed on fd {}: {}", what, fd, strerror(err));
        throw std::runtime_error(std::string("Error setting socket option: ") + what);
    }
}

static int getFlags(int fd, const char* why)
{
    int flags = ::fcntl(fd, F_GETFL, 0);
    if (flags < 0) {
        SPDLOG_ERROR("fcntl(F_GETFL) failed on fd {}: {}", fd, strerror(errno));
        throw std::runtime_error(std::string("Error ") + why);
    }
    return flags;
}

void setReuseAddr(int fd)
{
    setOpt(fd, SOL_SOCKET, SO_REUSEADDR, 1, "SO_REUSEADDR");
}

void setNoDelay(int fd)
{
    setOpt(fd, IPPROTO_TCP, TCP_NODELAY, 1, "TCP_NODELAY");
}

void setQuickAck(int fd)
{
    setOpt(fd, IPPROTO_TCP, TCP_QUICKACK, 1, "TCP_QUICKACK", true);
}

void setBusyPolling(int fd)
{
    // Microseconds to busy poll in the kernel before sleeping
    setOpt(fd, SOL_SOCKET, SO_BUSY_POLL, 10000, "SO_BUSY_POLL", true);
}

void setNonBlocking(int fd)
{
    int flags = getFlags(fd, "setting socket as non-blocking");
    ::fcntl(fd, F_SETFL, flags | O_NONBLOCK);
}

void setBlocking(int fd)
{
    int flags = getFlags(fd, "setting socket as blocking");
    ::fcntl(fd, F_SETFL, flags & ~O_NONBLOCK);
}

bool isNonBlocking(int fd)
{
    return (getFlags(fd, "checking if socket is blocking") & O_NONBLOCK) != 0;
}

static void setTimeout(int fd, int name, int timeoutMs, const char* what)
{
    timeval tv;
    tv.tv_sec = timeoutMs / 1000;
    tv.tv_usec = (timeoutMs % 1000) * 1000;
    if (::setsockopt(fd, SOL_SOCKET, name, &tv, sizeof(tv)) != 0) {
        SPDLOG_ERROR("setsockopt {} failed on fd {}: {}", what, fd, strerror(errno));
        throw std::runtime_error(std::string("Error setting ") + what);
    }
}

void setRecvTimeoutMs(int fd, int timeoutMs)
{
    setTimeout(fd, SO_RCVTIMEO, timeoutMs, "recv timeout");
}

void setSendTimeoutMs(int fd, int timeoutMs)
{
    setTimeout(fd, SO_SNDTIMEO, timeoutMs, "send timeout");
}

void setRecvBufferSize(int fd, size_t bufferSize)
{
    setOpt(fd, SOL_SOCKET, SO_RCVBUF, (int)bufferSize, "SO_RCVBUF");
}

void setSendBufferSize(int fd, size_t bufferSize)
{
    setOpt(fd, SOL_SOCKET, SO_SNDBUF, (int)bufferSize, "SO_SNDBUF");
}

Socket::Socket()
{
    fd = ::socket(AF_INET, SOCK_STREAM, 0);
    if (fd < 0) {
        throw std::runtime_error(std::string("socket() failed: ") + strerror(errno));
    }
}

Socket::Socket(int connFd)
  : fd(connFd)
{}

Socket::Socket(Socket&& other) noexcept
  : fd(other.fd)
{
    other.fd = -1;
}

Socket& Socket::operator=(Socket&& other) noexcept
{
    if (this != &other) {
        close();
        fd = other.fd;
        other.fd = -1;
    }
    return *this;
}

Socket::~Socket()
{
    close();
}

void Socket::close()
{
    if (fd >= 0) {
        ::close(fd);
        fd = -1;
    }
}

static sockaddr_in resolve(const std::string& host, int port)
{
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    if (::inet_pton(AF_INET, host.c_str(), &addr.sin_addr) == 1) {
        return addr;
    }
    addrinfo hints;
    memset(&hints, 0, sizeof(hints));
    hints.ai_family = AF_INET;
    hints.ai_socktype = SOCK_STREAM;
    addrinfo* res = nullptr;
    if (::getaddrinfo(host.c_str(), nullptr, &hints, &res) != 0 || res == nullptr) {
        throw std::runtime_error("Could not resolve host " + host);
    }
    addr.sin_addr = ((sockaddr_in*)res->ai_addr)->sin_addr;
    ::freeaddrinfo(res);
    return addr;
}

SendSocket::SendSocket(const std::string& hostIn, int portIn)
  : host(hostIn)
  , port(portIn)
{}

void SendSocket::dial(int retries, int sleepMs)
{
    sockaddr_in addr = resolve(host, port);
    for (int attempt = 0; attempt < retries; attempt++) {
        if (::connect(sock.get(), (sockaddr*)&addr, sizeof(addr)) == 0) {
            setNoDelay(sock.get());
            setQuickAck(sock.get());
            setSendBufferSize(sock.get(), SocketBufferSizeBytes);
            return;
        }
        SPDLOG_TRACE("Retrying connection to {}:{} ({})", host, port, strerror(errno));
        // A failed connect leaves the socket in an unspecified state
        sock = Socket();
        std::this_thread::sleep_for(std::chrono::milliseconds(sleepMs));
    }
    SPDLOG_ERROR("Error connecting to {}:{}: {}", host, port, strerror(errno));
    throw std::runtime_error("Error connecting to remote TCP socket");
}

void SendSocket::sendOne(const uint8_t* buffer, size_t bufferSize)
{
    size_t sent = 0;
    while (sent < bufferSize) {
        ssize_t n = ::send(sock.get(), buffer + sent, bufferSize - sent, MSG_NOSIGNAL);
        if (n < 0) {
            if (errno == EINTR) {
                continue;
            }
            if (errno == EAGAIN || errno == EWOULDBLOCK) {
                CPU_RELAX();
                continue;
            }
            SPDLOG_ERROR("TCP send to {}:{} failed: {}", host, port, strerror(errno));
            throw std::runtime_error("Error sending TCP message");
        }
        sent += (size_t)n;
    }
}

Address::Address(const std::string& host, int port)
{
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    if (::inet_pton(AF_INET, host.c_str(), &addr.sin_addr) != 1) {
        throw std::runtime_error("Not an IPv4 address: " + host);
    }
}

Address::Address(int port)
{
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    addr.sin_addr.s_addr = htonl(INADDR_ANY);
}

int Address::port() const
{
    return ntohs(addr.sin_port);
}

std::string Address::host() const
{
    char buf[INET_ADDRSTRLEN] = { 0 };
    ::inet_ntop(AF_INET, &addr.sin_addr, buf, sizeof(buf));
    return buf;
}

RecvSocket::RecvSocket(int portIn, const std::string& hostIn)
  : host(hostIn)
  , port(portIn)
{}

RecvSocket::~RecvSocket()
{
    for (int c : openConnections) {
        ::close(c);
    }
}

void RecvSocket::listen()
{
    setReuseAddr(sock.get());
    sockaddr_in addr;
    memset(&addr, 0, sizeof(addr));
    addr.sin_family = AF_INET;
    addr.sin_port = htons((uint16_t)port);
    ::inet_pton(AF_INET, host.c_str(), &addr.sin_addr);
    int rc = -1;
    for (int attempt = 0; attempt < 5; attempt++) {
        rc = ::bind(sock.get(), (sockaddr*)&addr, sizeof(addr));
        if (rc == 0) {
            break;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(200));
    }
    if (rc != 0) {
        SPDLOG_ERROR("Error binding to {}:{}: {}", host, port, strerror(errno));
        throw std::runtime_error("Socket error binding to port");
    }
    if (port == 0) {
        socklen_t len = sizeof(addr);
        ::getsockname(sock.get(), (sockaddr*)&addr, &len);
        port = ntohs(addr.sin_port);
    }
    if (::listen(sock.get(), SocketListenBacklog) != 0) {
        throw std::runtime_error("Socket error listening");
    }
}

int RecvSocket::accept(int timeoutMs)
{
    pollfd p{ sock.get(), POLLIN, 0 };
    int pr = ::poll(&p, 1, timeoutMs);
    if (pr <= 0) {
        SPDLOG_ERROR("Timed out accepting on port {}", port);
        throw std::runtime_error("Time-out polling for accept");
    }
    int conn = ::accept(sock.get(), nullptr, nullptr);
    if (conn < 0) {
        throw std::runtime_error(std::string("Error accepting connection: ") + strerror(errno));
    }
    setNoDelay(conn);
    setQuickAck(conn);
    setRecvBufferSize(conn, SocketBufferSizeBytes);
    openConnections.push_back(conn);
    return conn;
}

void RecvSocket::recvOne(int conn, uint8_t* buffer, size_t bufferSize)
{
    size_t got = 0;
    while (got < bufferSize) {
        ssize_t n = ::recv(conn, buffer + got, bufferSize - got, 0);
        if (n == 0) {
            throw std::runtime_error("TCP connection closed by peer");
        }
        if (n < 0) {
            if (errno == EINTR) {
                continue;
            }
            if (errno == EAGAIN || errno == EWOULDBLOCK) {
                // A non-blocking socket is polled; on a blocking one this is
                // its receive timeout expiring
                if ((::fcntl(conn, F_GETFL, 0) & O_NONBLOCK) != 0) {
                    CPU_RELAX();
                    continue;
                }
                SPDLOG_ERROR("TCP recv on fd {} timed out", conn);
                throw std::runtime_error("TCP receive timed out");
            }
            SPDLOG_ERROR("TCP recv failed: {}", strerror(errno));
            throw std::runtime_error("Error receiving TCP message");
        }
        got += (size_t)n;
    }
}

} // namespace faabric::transport::tcp
