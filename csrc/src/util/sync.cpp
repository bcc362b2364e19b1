// FlagWaiter, Latch, Barrier, PeriodicBackgroundThread, TokenPool
#include <faabric/util/PeriodicBackgroundThread.h>
#include <faabric/util/barrier.h>
#include <faabric/util/latch.h>
#include <faabric/util/locks.h>
#include <faabric/util/logging.h>
#include <faabric/util/queue.h>

#include <chrono>
#include <stdexcept>

namespace faabric::util {

FlagWaiter::FlagWaiter(int timeoutMsIn)
  : timeoutMs(timeoutMsIn)
{}

void FlagWaiter::waitOnFlag()
{
    // Keep ourselves alive in case the owner drops its reference meanwhile
    auto self = weak_from_this().lock();
    if (flag.load()) {
        return;
    }
    UniqueLock lock(flagMx);
    if (!cv.wait_for(lock, std::chrono::milliseconds(timeoutMs), [this] {
            return flag.load();
        })) {
        SPDLOG_ERROR("Timed out waiting for flag");
        throw std::runtime_error("Timed out waiting for flag");
    }
}

void FlagWaiter::setFlag(bool value)
{
    {
        UniqueLock lock(flagMx);
        flag.store(value);
    }
    cv.notify_all();
}

std::shared_ptr<Latch> Latch::create(int count, int timeoutMs)
{
    return std::make_shared<Latch>(count, timeoutMs);
}

Latch::Latch(int countIn, int timeoutMsIn)
  : count(countIn)
  , timeoutMs(timeoutMsIn)
{}

void Latch::wait()
{
    UniqueLock lock(mx);
    waiters++;
    if (waiters > count) {
        throw std::runtime_error("Latch already used");
    }
    if (waiters == count) {
        cv.notify_all();
        return;
    }
    if (!cv.wait_for(lock, std::chrono::milliseconds(timeoutMs), [this] {
            return waiters >= count;
        })) {
        throw std::runtime_error("Latch timed out");
    }
}

std::shared_ptr<Barrier> Barrier::create(int count,
                                         std::function<void()> completionFunction,
                                         int timeoutMs)
{
    return std::make_shared<Barrier>(
      count, std::move(completionFunction), timeoutMs);
}

Barrier::Barrier(int countIn,
                 std::function<void()> completionFunctionIn,
                 int timeoutMsIn)
  : count(countIn)
  , completionFunction(std::move(completionFunctionIn))
  , timeoutMs(timeoutMsIn)
{}

void Barrier::wait()
{
    UniqueLock lock(mx);
    uint64_t gen = generation;
    arrived++;
    if (arrived == count) {
        completionFunction();
        arrived = 0;
        generation++;
        cv.notify_all();
        return;
    }
    if (!cv.wait_for(lock, std::chrono::milliseconds(timeoutMs), [this, gen] {
            return generation != gen;
        })) {
        throw std::runtime_error("Barrier timed out");
    }
}

PeriodicBackgroundThread::~PeriodicBackgroundThread()
{
    stop();
}

void PeriodicBackgroundThread::start(int intervalSecondsIn)
{
    startMs(intervalSecondsIn * 1000);
}

void PeriodicBackgroundThread::startMs(int intervalMsIn)
{
    stop();
    intervalMs = intervalMsIn;
    if (intervalMs <= 0) {
        SPDLOG_DEBUG("Periodic thread disabled (interval {}ms)", intervalMs);
        return;
    }
    workThread = std::make_unique<std::jthread>([this](std::stop_token st) {
        while (!st.stop_requested()) {
            {
                std::unique_lock<std::mutex> lock(mx);
                bool stopped = timeoutCv.wait_for(
                  lock, st, std::chrono::milliseconds(intervalMs), [&st] {
                      return st.stop_requested();
                  });
                if (stopped) {
                    break;
                }
            }
            try {
                doWork();
            } catch (const std::exception& e) {
                SPDLOG_ERROR("Periodic work failed: {}", e.what());
            }
        }
        tidyUp();
    });
}

void PeriodicBackgroundThread::stop()
{
    if (workThread == nullptr) {
        return;
    }
    workThread->request_stop();
    timeoutCv.notify_all();
    if (workThread->joinable()) {
        workThread->join();
    }
    workThread.reset();
}

void PeriodicBackgroundThread::tidyUp() {}

TokenPool::TokenPool(int nTokens)
  : _size(nTokens)
{
    for (int i = 0; i < nTokens; i++) {
        queue.enqueue(i);
    }
}

int TokenPool::getToken()
{
    if (_size == 0) {
        return -1;
    }
    return queue.dequeue();
}

void TokenPool::releaseToken(int token)
{
    queue.enqueue(token);
}

void TokenPool::reset()
{
    queue.reset();
    for (int i = 0; i < _size; i++) {
        queue.enqueue(i);
    }
}

int TokenPool::size()
{
    return _size;
}

int TokenPool::taken()
{
    return _size - (int)queue.size();
}

int TokenPool::free()
{
    return (int)queue.size();
}

} // namespace faabric::util
