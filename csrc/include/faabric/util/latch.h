#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>

#define DEFAULT_LATCH_TIMEOUT_MS 10000

namespace faabric::util {

// Count-down latch where every participant calls wait() exactly once
class Latch
{
  public:
    static std::shared_ptr<Latch> create(
      int count,
      int timeoutMs = DEFAULT_LATCH_TIMEOUT_MS);

    explicit Latch(int countIn, int timeoutMsIn = DEFAULT_LATCH_TIMEOUT_MS);

    // Throws if more than `count` callers arrive, or on timeout
    void wait();

  private:
    int count;
    int waiters = 0;
    int timeoutMs;
    std::mutex mx;
    std::condition_variable cv;
};

}
