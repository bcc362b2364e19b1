// Out-of-band bootstrap between the per-GPU processes of one box: a star of
// abstract-namespace Unix sockets rooted at rank 0.  Gives the device layer the
// three primitives it needs to wire up peer memory across processes:
// allgather of small blobs (IPC handles), barrier, and file-descriptor exchange
// (SCM_RIGHTS) for VMM / multicast shareable handles.  Control plane only - no
// payload ever travels here.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace faabric::device {

class Bootstrap
{
  public:
    // jobId must be identical on all ranks (e.g. MASTER_PORT of the launcher)
    Bootstrap(int rank, int nranks, const std::string& jobId, int timeoutMs = 60000);
    ~Bootstrap();

    Bootstrap(const Bootstrap&) = delete;
    Bootstrap& operator=(const Bootstrap&) = delete;

    int rank() const { return rank_; }
    int size() const { return nranks_; }

    // Every rank contributes `len` bytes; returns nranks*len bytes rank-ordered
    std::vector<uint8_t> allGather(const void* data, size_t len);

    void barrier();

    // Every rank contributes one fd; returns one (duplicated) fd per rank.
    // Entry [rank] is a dup of the caller's own fd.  Caller closes them.
    std::vector<int> allGatherFds(int myFd);

    // Rank `root` contributes an fd, everybody receives a duplicate
    int broadcastFd(int fd, int root);

  private:
    int rank_;
    int nranks_;
    int timeoutMs_;
    int listenFd_ = -1;
    // root: one connection per peer rank (index = rank); others: [0] = root
    std::vector<int> conns_;

    void sendAll(int fd, const void* buf, size_t len);
    void recvAll(int fd, void* buf, size_t len);
    void sendFd(int sock, int fd);
    int recvFd(int sock);
};

} // namespace faabric::device
