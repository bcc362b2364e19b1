#include "faabric/device/bootstrap.h"

#include <cerrno>
#include <chrono>
#include <cstring>
#include <poll.h>
#include <stdexcept>
#include <sys/socket.h>
#include <sys/un.h>
#include <thread>
#include <unistd.h>

namespace faabric::device {

static sockaddr_un makeAddr(const std::string& jobId, socklen_t& len)
{
    sockaddr_un addr;
    memset(&addr, 0, sizeof(addr));
    addr.sun_family = AF_UNIX;
    std::string name = "faabric-b200-" + jobId;
    if (name.size() > sizeof(addr.sun_path) - 2) {
        name.resize(sizeof(addr.sun_path) - 2);
    }
    // abstract namespace: leading NUL, no filesystem entry to clean up
    memcpy(addr.sun_path + 1, name.data(), name.size());
    len = (socklen_t)(offsetof(sockaddr_un, sun_path) + 1 + name.size());
    return addr;
}

Bootstrap::Bootstrap(int rank,
                     int nranks,
                     const std::string& jobId,
                     int timeoutMs)
  : rank_(rank)
  , nranks_(nranks)
  , timeoutMs_(timeoutMs)
{
    if (nranks_ <= 1) {
        return;
    }
    socklen_t alen;
    sockaddr_un addr = makeAddr(jobId, alen);
    if (rank_ == 0) {
        listenFd_ = ::socket(AF_UNIX, SOCK_STREAM, 0);
        if (listenFd_ < 0) {
            throw std::runtime_error("bootstrap: socket() failed");
        }
        if (::bind(listenFd_, (sockaddr*)&addr, alen) != 0) {
            throw std::runtime_error(std::string("bootstrap: bind failed: ") +
                                     strerror(errno));
        }
        ::listen(listenFd_, nranks_);
        conns_.assign(nranks_, -1);
        for (int i = 1; i < nranks_; i++) {
            pollfd p{ listenFd_, POLLIN, 0 };
            int pr = ::poll(&p, 1, timeoutMs_);
            if (pr <= 0) {
                throw std::runtime_error(
                  "bootstrap: timed out waiting for peers");
            }
            int c = ::accept(listenFd_, nullptr, nullptr);
            if (c < 0) {
                throw std::runtime_error("bootstrap: accept failed");
            }
            int32_t peerRank = -1;
            recvAll(c, &peerRank, sizeof(peerRank));
            if (peerRank <= 0 || peerRank >= nranks_ ||
                conns_[peerRank] != -1) {
                throw std::runtime_error("bootstrap: bad peer rank");
            }
            conns_[peerRank] = c;
        }
    } else {
        int c = -1;
        auto deadline = std::chrono::steady_clock::now() +
                        std::chrono::milliseconds(timeoutMs_);
        while (true) {
            c = ::socket(AF_UNIX, SOCK_STREAM, 0);
            if (::connect(c, (sockaddr*)&addr, alen) == 0) {
                break;
            }
            ::close(c);
            if (std::chrono::steady_clock::now() > deadline) {
                throw std::runtime_error(
                  "bootstrap: timed out connecting to rank 0");
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(20));
        }
        int32_t r = rank_;
        sendAll(c, &r, sizeof(r));
        conns_.assign(1, c);
    }
}

Bootstrap::~Bootstrap()
{
    for (int c : conns_) {
        if (c >= 0) {
            ::close(c);
        }
    }
    if (listenFd_ >= 0) {
        ::close(listenFd_);
    }
}

void Bootstrap::sendAll(int fd, const void* buf, size_t len)
{
    const uint8_t* p = (const uint8_t*)buf;
    while (len > 0) {
        ssize_t n = ::send(fd, p, len, MSG_NOSIGNAL);
        if (n <= 0) {
            if (n < 0 && errno == EINTR) {
                continue;
            }
            throw std::runtime_error("bootstrap: send failed");
        }
        p += n;
        len -= (size_t)n;
    }
}

void Bootstrap::recvAll(int fd, void* buf, size_t len)
{
    uint8_t* p = (uint8_t*)buf;
    while (len > 0) {
        pollfd pf{ fd, POLLIN, 0 };
        int pr = ::poll(&pf, 1, timeoutMs_);
        if (pr == 0) {
            throw std::runtime_error("bootstrap: recv timed out");
        }
        ssize_t n = ::recv(fd, p, len, 0);
        if (n <= 0) {
            if (n < 0 && errno == EINTR) {
                continue;
            }
            throw std::runtime_error("bootstrap: recv failed / peer closed");
        }
        p += n;
        len -= (size_t)n;
    }
}

std::vector<uint8_t> Bootstrap::allGather(const void* data, size_t len)
{
    std::vector<uint8_t> out((size_t)nranks_ * len);
    if (nranks_ <= 1) {
        memcpy(out.data(), data, len);
        return out;
    }
    if (rank_ == 0) {
        memcpy(out.data(), data, len);
        for (int r = 1; r < nranks_; r++) {
            recvAll(conns_[r], out.data() + (size_t)r * len, len);
        }
        for (int r = 1; r < nranks_; r++) {
            sendAll(conns_[r], out.data(), out.size());
        }
    } else {
        sendAll(conns_[0], data, len);
        recvAll(conns_[0], out.data(), out.size());
    }
    return out;
}

void Bootstrap::barrier()
{
    uint8_t b = 1;
    allGather(&b, 1);
}

void Bootstrap::sendFd(int sock, int fd)
{
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    char payload = 'F';
    iovec io{ &payload, 1 };
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    cmsghdr* cm = CMSG_FIRSTHDR(&msg);
    cm->cmsg_level = SOL_SOCKET;
    cm->cmsg_type = SCM_RIGHTS;
    cm->cmsg_len = CMSG_LEN(sizeof(int));
    memcpy(CMSG_DATA(cm), &fd, sizeof(int));
    if (::sendmsg(sock, &msg, MSG_NOSIGNAL) != 1) {
        throw std::runtime_error("bootstrap: sendmsg(SCM_RIGHTS) failed");
    }
}

int Bootstrap::recvFd(int sock)
{
    pollfd pf{ sock, POLLIN, 0 };
    if (::poll(&pf, 1, timeoutMs_) <= 0) {
        throw std::runtime_error("bootstrap: fd recv timed out");
    }
    msghdr msg;
    memset(&msg, 0, sizeof(msg));
    char payload = 0;
    iovec io{ &payload, 1 };
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    char ctrl[CMSG_SPACE(sizeof(int))];
    memset(ctrl, 0, sizeof(ctrl));
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    if (::recvmsg(sock, &msg, 0) != 1) {
        throw std::runtime_error("bootstrap: recvmsg failed");
    }
    cmsghdr* cm = CMSG_FIRSTHDR(&msg);
    if (cm == nullptr || cm->cmsg_type != SCM_RIGHTS) {
        throw std::runtime_error("bootstrap: no fd in message");
    }
    int fd = -1;
    memcpy(&fd, CMSG_DATA(cm), sizeof(int));
    return fd;
}

std::vector<int> Bootstrap::allGatherFds(int myFd)
{
    std::vector<int> fds(nranks_, -1);
    if (nranks_ <= 1) {
        fds[0] = ::dup(myFd);
        return fds;
    }
    if (rank_ == 0) {
        fds[0] = ::dup(myFd);
        for (int r = 1; r < nranks_; r++) {
            fds[r] = recvFd(conns_[r]);
        }
        for (int r = 1; r < nranks_; r++) {
            for (int k = 0; k < nranks_; k++) {
                sendFd(conns_[r], fds[k]);
            }
        }
    } else {
        sendFd(conns_[0], myFd);
        for (int k = 0; k < nranks_; k++) {
            fds[k] = recvFd(conns_[0]);
        }
    }
    return fds;
}

int Bootstrap::broadcastFd(int fd, int root)
{
    if (nranks_ <= 1) {
        return ::dup(fd);
    }
    // Route through rank 0
    if (rank_ == 0) {
        int src = fd;
        bool owned = false;
        if (root != 0) {
            src = recvFd(conns_[root]);
            owned = true;
        }
        for (int r = 1; r < nranks_; r++) {
            sendFd(conns_[r], src);
        }
        return owned ? src : ::dup(src);
    }
    if (rank_ == root) {
        sendFd(conns_[0], fd);
    }
    return recvFd(conns_[0]);
}

} // namespace faabric::device
