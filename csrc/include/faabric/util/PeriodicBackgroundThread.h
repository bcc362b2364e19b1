#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>

namespace faabric::util {

// Runs doWork() every `intervalSeconds` on its own thread until stop()
class PeriodicBackgroundThread
{
  public:
    virtual ~PeriodicBackgroundThread();

    void start(int intervalSecondsIn);

    // Millisecond resolution variant (used by tests and the keep-alive)
    void startMs(int intervalMsIn);

    void stop();

    virtual void doWork() = 0;

    int getIntervalSeconds() const { return intervalMs / 1000; }

    // Hook called once on the worker thread when it exits
    virtual void tidyUp();

  private:
    std::unique_ptr<std::jthread> workThread;
    std::mutex mx;
    std::condition_variable_any timeoutCv;
    int intervalMs = 0;
};

}
