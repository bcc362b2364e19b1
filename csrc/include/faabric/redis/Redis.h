// Redis-compatible store.  The reference talks to a real Redis over hiredis
// (include/faabric/redis/Redis.h:17-228); on a single box this is an
// in-process store with the same verbs: strings, ranges, sets, lists (with
// blocking dequeue), conditional delete and locks.  One instance per role
// (QUEUE / STATE), shared by all threads of the process.
#pragma once

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace faabric::redis {

enum RedisRole
{
    QUEUE,
    STATE,
};

class RedisNoResponseException : public std::runtime_error
{
  public:
    RedisNoResponseException()
      : std::runtime_error("No response from redis (timeout)")
    {}
};

class Redis
{
  public:
    ~Redis() = default;

    static Redis& getQueue();

    static Redis& getState();

    // ---- Standard Redis commands ----
    void ping();

    std::vector<uint8_t> get(const std::string& key);

    size_t strlen(const std::string& key);

    void get(const std::string& key, uint8_t* buffer, size_t size);

    void set(const std::string& key, const std::vector<uint8_t>& value);

    void set(const std::string& key, const uint8_t* value, size_t size);

    void del(const std::string& key);

    long getCounter(const std::string& key);

    long incr(const std::string& key);

    long decr(const std::string& key);

    long incrByLong(const std::string& key, long val);

    long decrByLong(const std::string& key, long val);

    void setRange(const std::string& key,
                  long offset,
                  const uint8_t* value,
                  size_t size);

    // Pipelined variant: queued then flushed (immediate here)
    void setRangePipeline(const std::string& key,
                          long offset,
                          const uint8_t* value,
                          size_t size);

    void flushPipeline(long pipelineLength);

    void getRange(const std::string& key,
                  uint8_t* buffer,
                  size_t bufferLen,
                  long start,
                  long end);

    void sadd(const std::string& key, const std::string& value);

    void srem(const std::string& key, const std::string& value);

    long scard(const std::string& key);

    bool sismember(const std::string& key, const std::string& value);

    std::string srandmember(const std::string& key);

    std::set<std::string> smembers(const std::string& key);

    std::set<std::string> sdiff(const std::string& keyA, const std::string& keyB);

    std::set<std::string> sinter(const std::string& keyA, const std::string& keyB);

    int lpushLong(const std::string& key, long value);

    int rpushLong(const std::string& key, long value);

    void flushAll();

    long listLength(const std::string& queueName);

    long getTtl(const std::string& key);

    void expire(const std::string& key, long expiry);

    void refresh();

    // ---- Locks ----
    uint32_t acquireLock(const std::string& key, int expirySeconds);

    void releaseLock(const std::string& key, uint32_t lockId);

    void delIfEq(const std::string& key, uint32_t value);

    bool setnxex(const std::string& key, long value, int expirySeconds);

    long getLong(const std::string& key);

    void setLong(const std::string& key, long value);

    // ---- Queueing ----
    void enqueue(const std::string& queueName, const std::string& value);

    void enqueueBytes(const std::string& queueName, const std::vector<uint8_t>& value);

    void enqueueBytes(const std::string& queueName, const uint8_t* buffer, size_t bufferLen);

    std::string dequeue(const std::string& queueName, int timeout = 60000);

    std::vector<uint8_t> dequeueBytes(const std::string& queueName, int timeout = 60000);

    void dequeueBytes(const std::string& queueName, uint8_t* buffer, size_t bufferLen, int timeout = 60000);

    void dequeueMultiple(const std::string& queueName, uint8_t* buff, long buffLen, long nElems);

    // ---- Scheduler notification ----
    void publishSchedulerResult(const std::string& key, const std::string& statusKey, const std::vector<uint8_t>& result);

  private:
    explicit Redis(RedisRole roleIn);

    RedisRole role;
    std::mutex mx;
    std::condition_variable listCv;
    std::unordered_map<std::string, std::vector<uint8_t>> strings;
    std::unordered_map<std::string, std::set<std::string>> sets;
    std::unordered_map<std::string, std::deque<std::vector<uint8_t>>> lists;
    std::unordered_map<std::string, long> expiries; // epoch ms
    uint32_t nextLockId = 1;

    bool isExpiredLocked(const std::string& key);

    std::vector<uint8_t> popFront(const std::string& queueName, int timeoutMs);
};

}
