#pragma once

#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>

#define DEFAULT_BARRIER_TIMEOUT_MS 10000

namespace faabric::util {

// Reusable (cyclic) thread barrier with a completion hook and timeout
class Barrier
{
  public:
    static std::shared_ptr<Barrier> create(
      int count,
      std::function<void()> completionFunction = []() {},
      int timeoutMs = DEFAULT_BARRIER_TIMEOUT_MS);

    explicit Barrier(int countIn,
                     std::function<void()> completionFunctionIn,
                     int timeoutMsIn);

    void wait();

  private:
    int count;
    int arrived = 0;
    uint64_t generation = 0;
    std::function<void()> completionFunction;
    int timeoutMs;
    std::mutex mx;
    std::condition_variable cv;
};

}
