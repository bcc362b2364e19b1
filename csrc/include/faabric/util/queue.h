// Host-side queues (reference: include/faabric/util/queue.h:24-265).
//  Queue<T>              mutex + condvar, timeouts, peek, drain
//  FixedCapacityQueue<T> bounded blocking SPSC/MPMC ring (own implementation)
//  SpinLockQueue<T>      bounded lock-free ring, busy-waiting (low latency)
//  TokenPool             pool of integer tokens
#pragma once

#include <thread>
#include <faabric/util/exception.h>
#include <faabric/util/locks.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <memory>
#include <optional>
#include <queue>
#include <set>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#define FAABRIC_CPU_PAUSE() _mm_pause()
#else
#define FAABRIC_CPU_PAUSE() std::this_thread::yield()
#endif

#define DEFAULT_QUEUE_TIMEOUT_MS 5000
#define DEFAULT_QUEUE_SIZE 1024

namespace faabric::util {

class QueueTimeoutException : public faabric::util::FaabricException
{
  public:
    explicit QueueTimeoutException(std::string message)
      : FaabricException(std::move(message))
    {}
};

template<typename T>
class Queue
{
  public:
    void enqueue(T value)
    {
        {
            UniqueLock lock(mx);
            mq.emplace(std::move(value));
            approxSize.store((long)mq.size(), std::memory_order_release);
        }
        enqueueNotifier.notify_one();
    }

    void dequeueIfPresent(T* res)
    {
        UniqueLock lock(mx);
        if (!mq.empty()) {
            T value = std::move(mq.front());
            mq.pop();
            approxSize.store((long)mq.size(), std::memory_order_release);
            emptyNotifier.notify_one();
            *res = std::move(value);
        }
    }

    T dequeue(long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        if (timeoutMs <= 0) {
            throw std::runtime_error("Dequeue timeout must be positive");
        }
        // A consumer in the middle of a request/response exchange gets its
        // next item within microseconds: look for it briefly before paying
        // for a sleep + wake-up (the yield lets a producer that shares our
        // core run)
        if (approxSize.load(std::memory_order_acquire) == 0) {
            auto start = std::chrono::steady_clock::now();
            for (int i = 0; approxSize.load(std::memory_order_acquire) == 0; i++) {
                if ((i & 15) == 15) {
                    std::this_thread::yield();
                    if (std::chrono::steady_clock::now() - start > std::chrono::microseconds(20)) {
                        break;
                    }
                }
            }
        }
        UniqueLock lock(mx);
        if (!enqueueNotifier.wait_for(lock,
                                      std::chrono::milliseconds(timeoutMs),
                                      [this] { return !mq.empty(); })) {
            throw QueueTimeoutException("Timeout waiting for dequeue");
        }
        T value = std::move(mq.front());
        mq.pop();
        approxSize.store((long)mq.size(), std::memory_order_release);
        emptyNotifier.notify_one();
        return value;
    }

    T* peek(long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        UniqueLock lock(mx);
        if (timeoutMs <= 0) {
            throw std::runtime_error("Peek timeout must be positive");
        }
        if (!enqueueNotifier.wait_for(lock,
                                      std::chrono::milliseconds(timeoutMs),
                                      [this] { return !mq.empty(); })) {
            throw QueueTimeoutException("Timeout waiting for queue to peek");
        }
        return &mq.front();
    }

    void waitToDrain(long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        UniqueLock lock(mx);
        if (!emptyNotifier.wait_for(lock,
                                    std::chrono::milliseconds(timeoutMs),
                                    [this] { return mq.empty(); })) {
            throw QueueTimeoutException("Timed out waiting for queue to empty");
        }
    }

    void drain()
    {
        UniqueLock lock(mx);
        while (!mq.empty()) {
            mq.pop();
        }
        approxSize.store(0, std::memory_order_release);
        emptyNotifier.notify_all();
    }

    long size()
    {
        UniqueLock lock(mx);
        return (long)mq.size();
    }

    void reset()
    {
        UniqueLock lock(mx);
        std::queue<T> empty;
        std::swap(mq, empty);
        approxSize.store(0, std::memory_order_release);
    }

  private:
    std::queue<T> mq;
    std::atomic<long> approxSize{ 0 };
    std::condition_variable enqueueNotifier;
    std::condition_variable emptyNotifier;
    std::mutex mx;
};

// Bounded ring shared by both fixed-capacity variants.  Multi-producer /
// multi-consumer safe (sequence number per cell).
template<typename T>
class BoundedRing
{
  public:
    explicit BoundedRing(size_t capacityIn)
    {
        cap = 1;
        while (cap < capacityIn) {
            cap <<= 1;
        }
        cells = std::make_unique<Cell[]>(cap);
        for (size_t i = 0; i < cap; i++) {
            cells[i].seq.store(i, std::memory_order_relaxed);
        }
    }

    bool tryPush(T&& v)
    {
        size_t pos = head.load(std::memory_order_relaxed);
        while (true) {
            Cell& c = cells[pos & (cap - 1)];
            size_t seq = c.seq.load(std::memory_order_acquire);
            intptr_t dif = (intptr_t)seq - (intptr_t)pos;
            if (dif == 0) {
                if (head.compare_exchange_weak(
                      pos, pos + 1, std::memory_order_relaxed)) {
                    c.value = std::move(v);
                    c.seq.store(pos + 1, std::memory_order_release);
                    return true;
                }
            } else if (dif < 0) {
                return false; // full
            } else {
                pos = head.load(std::memory_order_relaxed);
            }
        }
    }

    bool tryPop(T& out)
    {
        size_t pos = tail.load(std::memory_order_relaxed);
        while (true) {
            Cell& c = cells[pos & (cap - 1)];
            size_t seq = c.seq.load(std::memory_order_acquire);
            intptr_t dif = (intptr_t)seq - (intptr_t)(pos + 1);
            if (dif == 0) {
                if (tail.compare_exchange_weak(
                      pos, pos + 1, std::memory_order_relaxed)) {
                    out = std::move(c.value);
                    c.seq.store(pos + cap, std::memory_order_release);
                    return true;
                }
            } else if (dif < 0) {
                return false; // empty
            } else {
                pos = tail.load(std::memory_order_relaxed);
            }
        }
    }

    size_t sizeApprox() const
    {
        size_t h = head.load(std::memory_order_relaxed);
        size_t t = tail.load(std::memory_order_relaxed);
        return h >= t ? h - t : 0;
    }

    size_t capacity() const { return cap; }

  private:
    struct Cell
    {
        std::atomic<size_t> seq;
        T value;
    };
    size_t cap;
    std::unique_ptr<Cell[]> cells;
    alignas(64) std::atomic<size_t> head{ 0 };
    alignas(64) std::atomic<size_t> tail{ 0 };
};

// Blocking bounded queue: spins briefly then sleeps on a condition variable
template<typename T>
class FixedCapacityQueue
{
  public:
    explicit FixedCapacityQueue(int capacity)
      : ring(capacity)
    {}

    FixedCapacityQueue()
      : ring(DEFAULT_QUEUE_SIZE)
    {}

    void enqueue(T value, long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        if (timeoutMs <= 0) {
            throw std::runtime_error("Enqueue timeout must be positive");
        }
        auto deadline = std::chrono::steady_clock::now() +
                        std::chrono::milliseconds(timeoutMs);
        while (!ring.tryPush(std::move(value))) {
            if (std::chrono::steady_clock::now() > deadline) {
                throw QueueTimeoutException("Timeout waiting for enqueue");
            }
            UniqueLock lock(mx);
            blockedProducers.fetch_add(1, std::memory_order_acq_rel);
            notFull.wait_for(lock, std::chrono::microseconds(200));
            blockedProducers.fetch_sub(1, std::memory_order_acq_rel);
        }
        // seq_cst pairing with the consumer: it registers as a sleeper
        // (under mx) BEFORE its final emptiness check
        std::atomic_thread_fence(std::memory_order_seq_cst);
        if (sleepers.load(std::memory_order_seq_cst) > 0) {
            UniqueLock lock(mx);
            notEmpty.notify_one();
        }
    }

    void dequeueIfPresent(T* res)
    {
        T v;
        if (ring.tryPop(v)) {
            *res = std::move(v);
            wakeProducer();
        }
    }

    T dequeue(long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        if (timeoutMs <= 0) {
            throw std::runtime_error("Dequeue timeout must be positive");
        }
        T v;
        // Phase 1: spin for a few tens of microseconds - a peer in the
        // middle of a ping-pong answers within that time
        auto start = std::chrono::steady_clock::now();
        for (int i = 0;; i++) {
            if (ring.tryPop(v)) {
                wakeProducer();
                return v;
            }
            FAABRIC_CPU_PAUSE();
            if ((i & 31) == 31) {
                // The producer may have been woken onto OUR core (wake-affine
                // placement): give it a chance instead of starving it
                std::this_thread::yield();
                if (std::chrono::steady_clock::now() - start > std::chrono::microseconds(SPIN_BEFORE_SLEEP_US)) {
                    break;
                }
            }
        }
        // Phase 2: sleep.  The emptiness check and the wait happen under the
        // same mutex the producer takes to notify, so no wake-up is lost
        auto deadline = start + std::chrono::milliseconds(timeoutMs);
        UniqueLock lock(mx);
        sleepers.fetch_add(1, std::memory_order_seq_cst);
        std::atomic_thread_fence(std::memory_order_seq_cst);
        while (true) {
            if (ring.tryPop(v)) {
                sleepers.fetch_sub(1, std::memory_order_acq_rel);
                lock.unlock();
                wakeProducer();
                return v;
            }
            if (std::chrono::steady_clock::now() > deadline) {
                sleepers.fetch_sub(1, std::memory_order_acq_rel);
                throw QueueTimeoutException("Timeout waiting for dequeue");
            }
            notEmpty.wait_for(lock, std::chrono::milliseconds(50));
        }
    }

    void drain()
    {
        T v;
        while (ring.tryPop(v)) {
        }
    }

    long size() { return (long)ring.sizeApprox(); }

    void reset() { drain(); }

  private:
    static constexpr int SPIN_BEFORE_SLEEP_US = 50;

    BoundedRing<T> ring;
    std::mutex mx;
    std::condition_variable notEmpty;
    std::condition_variable notFull;
    std::atomic<int> sleepers{ 0 };
    std::atomic<int> blockedProducers{ 0 };

    void wakeProducer()
    {
        if (blockedProducers.load(std::memory_order_acquire) > 0) {
            UniqueLock lock(mx);
            notFull.notify_one();
        }
    }
};

// Busy-waiting bounded queue for pinned rank threads
template<typename T>
class SpinLockQueue
{
  public:
    SpinLockQueue()
      : ring(DEFAULT_QUEUE_SIZE)
    {}

    explicit SpinLockQueue(int capacity)
      : ring(capacity)
    {}

    void enqueue(T& value, long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        T v = value;
        spinPush(std::move(v), timeoutMs);
    }

    void enqueue(T&& value, long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        spinPush(std::move(value), timeoutMs);
    }

    T dequeue(long timeoutMs = DEFAULT_QUEUE_TIMEOUT_MS)
    {
        T v;
        uint64_t spins = 0;
        auto start = std::chrono::steady_clock::now();
        while (!ring.tryPop(v)) {
            FAABRIC_CPU_PAUSE();
            if ((++spins & 0xffff) == 0 &&
                std::chrono::steady_clock::now() - start >
                  std::chrono::milliseconds(timeoutMs)) {
                throw QueueTimeoutException("Timeout spinning for dequeue");
            }
        }
        return v;
    }

    bool tryDequeue(T& out) { return ring.tryPop(out); }

    long size() { return (long)ring.sizeApprox(); }

    void drain()
    {
        T v;
        while (ring.tryPop(v)) {
        }
    }

    void reset() { drain(); }

  private:
    BoundedRing<T> ring;

    void spinPush(T&& v, long timeoutMs)
    {
        uint64_t spins = 0;
        auto start = std::chrono::steady_clock::now();
        while (!ring.tryPush(std::move(v))) {
            FAABRIC_CPU_PAUSE();
            if ((++spins & 0xffff) == 0 &&
                std::chrono::steady_clock::now() - start >
                  std::chrono::milliseconds(timeoutMs)) {
                throw QueueTimeoutException("Timeout spinning for enqueue");
            }
        }
    }
};

class TokenPool
{
  public:
    explicit TokenPool(int nTokens);

    int getToken();

    void releaseToken(int token);

    void reset();

    int size();

    int taken();

    int free();

  private:
    int _size;
    Queue<int> queue;
};

} // namespace faabric::util
