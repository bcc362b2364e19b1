#pragma once

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>

#define DEFAULT_FLAG_WAIT_MS 10000

namespace faabric::util {

typedef std::unique_lock<std::mutex> UniqueLock;
typedef std::unique_lock<std::shared_mutex> FullLock;
typedef std::shared_lock<std::shared_mutex> SharedLock;

// One-shot flag several threads can block on (with timeout)
class FlagWaiter : public std::enable_shared_from_this<FlagWaiter>
{
  public:
    explicit FlagWaiter(int timeoutMsIn = DEFAULT_FLAG_WAIT_MS);

    // Throws std::runtime_error on timeout
    void waitOnFlag();

    void setFlag(bool value);

  private:
    int timeoutMs;
    std::mutex flagMx;
    std::condition_variable cv;
    std::atomic<bool> flag = false;
};

}
