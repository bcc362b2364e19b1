#include <faabric/redis/Redis.h>
#include <faabric/util/clock.h>
#include <faabric/util/random.h>

#include <algorithm>
#include <cstring>

namespace faabric::redis {

Redis::Redis(RedisRole roleIn)
  : role(roleIn)
{}

Redis& Redis::getQueue()
{
    static Redis r(QUEUE);
    return r;
}

Redis& Redis::getState()
{
    static Redis r(STATE);
    return r;
}

void Redis::ping() {}

bool Redis::isExpiredLocked(const std::string& key)
{
    auto it = expiries.find(key);
    if (it == expiries.end()) {
        return false;
    }
    if (faabric::util::getGlobalClock().epochMillis() >= it->second) {
        expiries.erase(it);
        strings.erase(key);
        sets.erase(key);
        lists.erase(key);
        return true;
    }
    return false;
}

std::vector<uint8_t> Redis::get(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    isExpiredLocked(key);
    auto it = strings.find(key);
    return it == strings.end() ? std::vector<uint8_t>() : it->second;
}

size_t Redis::strlen(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    isExpiredLocked(key);
    auto it = strings.find(key);
    return it == strings.end() ? 0 : it->second.size();
}

void Redis::get(const std::string& key, uint8_t* buffer, size_t size)
{
    std::lock_guard<std::mutex> lk(mx);
    isExpiredLocked(key);
    auto it = strings.find(key);
    if (it == strings.end()) {
        return;
    }
    memcpy(buffer, it->second.data(), std::min(size, it->second.size()));
}

void Redis::set(const std::string& key, const std::vector<uint8_t>& value)
{
    set(key, value.data(), value.size());
}

void Redis::set(const std::string& key, const uint8_t* value, size_t size)
{
    std::lock_guard<std::mutex> lk(mx);
    strings[key].assign(value, value + size);
    expiries.erase(key);
}

void Redis::del(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    strings.erase(key);
    sets.erase(key);
    lists.erase(key);
    expiries.erase(key);
}

static long parseLong(const std::vector<uint8_t>& v)
{
    if (v.empty()) {
        return 0;
    }
    return strtol(std::string(v.begin(), v.end()).c_str(), nullptr, 10);
}

static std::vector<uint8_t> longBytes(long v)
{
    std::string s = std::to_string(v);
    return std::vector<uint8_t>(s.begin(), s.end());
}

long Redis::getCounter(const std::string& key)
{
    return parseLong(get(key));
}

long Redis::incrByLong(const std::string& key, long val)
{
    std::lock_guard<std::mutex> lk(mx);
    long v = parseLong(strings[key]) + val;
    strings[key] = longBytes(v);
    return v;
}

long Redis::decrByLong(const std::string& key, long val)
{
    return incrByLong(key, -val);
}

long Redis::incr(const std::string& key)
{
    return incrByLong(key, 1);
}

long Redis::decr(const std::string& key)
{
    return incrByLong(key, -1);
}

void Redis::setRange(const std::string& key, long offset, const uint8_t* value, size_t size)
{
    std::lock_guard<std::mutex> lk(mx);
    auto& v = strings[key];
    if (v.size() < (size_t)offset + size) {
        v.resize((size_t)offset + size, 0);
    }
    memcpy(v.data() + offset, value, size);
}

void Redis::setRangePipeline(const std::string& key, long offset, const uint8_t* value, size_t size)
{
    setRange(key, offset, value, size);
}

void Redis::flushPipeline(long pipelineLength) {}

void Redis::getRange(const std::string& key, uint8_t* buffer, size_t bufferLen, long start, long end)
{
    // Inclusive range like GETRANGE
    size_t rangeLen = (size_t)(end - start + 1);
    if (rangeLen > bufferLen) {
        throw std::runtime_error("Range " + std::to_string(start) + "-" + std::to_string(end) + " too long for buffer length " + std::to_string(bufferLen));
    }
    std::lock_guard<std::mutex> lk(mx);
    auto it = strings.find(key);
    if (it == strings.end() || (size_t)start >= it->second.size()) {
        return;
    }
    size_t n = std::min(rangeLen, it->second.size() - (size_t)start);
    memcpy(buffer, it->second.data() + start, n);
}

void Redis::sadd(const std::string& key, const std::string& value)
{
    std::lock_guard<std::mutex> lk(mx);
    sets[key].insert(value);
}

void Redis::srem(const std::string& key, const std::string& value)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = sets.find(key);
    if (it != sets.end()) {
        it->second.erase(value);
    }
}

long Redis::scard(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = sets.find(key);
    return it == sets.end() ? 0 : (long)it->second.size();
}

bool Redis::sismember(const std::string& key, const std::string& value)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = sets.find(key);
    return it != sets.end() && it->second.count(value) > 0;
}

std::string Redis::srandmember(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = sets.find(key);
    if (it == sets.end() || it->second.empty()) {
        return "";
    }
    int idx = faabric::util::randomInteger(0, (int)it->second.size() - 1);
    auto e = it->second.begin();
    std::advance(e, idx);
    return *e;
}

std::set<std::string> Redis::smembers(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = sets.find(key);
    return it == sets.end() ? std::set<std::string>() : it->second;
}

std::set<std::string> Redis::sdiff(const std::string& keyA, const std::string& keyB)
{
    std::set<std::string> a = smembers(keyA);
    std::set<std::string> b = smembers(keyB);
    std::set<std::string> out;
    std::set_difference(a.begin(), a.end(), b.begin(), b.end(), std::inserter(out, out.begin()));
    return out;
}

std::set<std::string> Redis::sinter(const std::string& keyA, const std::string& keyB)
{
    std::set<std::string> a = smembers(keyA);
    std::set<std::string> b = smembers(keyB);
    std::set<std::string> out;
    std::set_intersection(a.begin(), a.end(), b.begin(), b.end(), std::inserter(out, out.begin()));
    return out;
}

int Redis::lpushLong(const std::string& key, long value)
{
    std::lock_guard<std::mutex> lk(mx);
    lists[key].push_front(longBytes(value));
    listCv.notify_all();
    return (int)lists[key].size();
}

int Redis::rpushLong(const std::string& key, long value)
{
    std::lock_guard<std::mutex> lk(mx);
    lists[key].push_back(longBytes(value));
    listCv.notify_all();
    return (int)lists[key].size();
}

void Redis::flushAll()
{
    std::lock_guard<std::mutex> lk(mx);
    strings.clear();
    sets.clear();
    lists.clear();
    expiries.clear();
}

long Redis::listLength(const std::string& queueName)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = lists.find(queueName);
    return it == lists.end() ? 0 : (long)it->second.size();
}

long Redis::getTtl(const std::string& key)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = expiries.find(key);
    if (it == expiries.end()) {
        bool exists = strings.count(key) || sets.count(key) || lists.count(key);
        return exists ? -1 : -2;
    }
    long ms = it->second - faabric::util::getGlobalClock().epochMillis();
    return std::max<long>(0, ms / 1000);
}

void Redis::expire(const std::string& key, long expiry)
{
    std::lock_guard<std::mutex> lk(mx);
    expiries[key] = faabric::util::getGlobalClock().epochMillis() + expiry * 1000;
}

void Redis::refresh() {}

bool Redis::setnxex(const std::string& key, long value, int expirySeconds)
{
    std::lock_guard<std::mutex> lk(mx);
    isExpiredLocked(key);
    if (strings.count(key) > 0) {
        return false;
    }
    strings[key] = longBytes(value);
    expiries[key] = faabric::util::getGlobalClock().epochMillis() + (long)expirySeconds * 1000;
    return true;
}

uint32_t Redis::acquireLock(const std::string& key, int expirySeconds)
{
    uint32_t id;
    {
        std::lock_guard<std::mutex> lk(mx);
        id = nextLockId++;
    }
    std::string lockKey = key + "_lock";
    return setnxex(lockKey, id, expirySeconds) ? id : 0;
}

void Redis::releaseLock(const std::string& key, uint32_t lockId)
{
    delIfEq(key + "_lock", lockId);
}

void Redis::delIfEq(const std::string& key, uint32_t value)
{
    std::lock_guard<std::mutex> lk(mx);
    auto it = strings.find(key);
    if (it != strings.end() && parseLong(it->second) == (long)value) {
        strings.erase(it);
        expiries.erase(key);
    }
}

long Redis::getLong(const std::string& key)
{
    return parseLong(get(key));
}

void Redis::setLong(const std::string& key, long value)
{
    set(key, longBytes(value));
}

void Redis::enqueue(const std::string& queueName, const std::string& value)
{
    enqueueBytes(queueName, (const uint8_t*)value.data(), value.size());
}

void Redis::enqueueBytes(const std::string& queueName, const std::vector<uint8_t>& value)
{
    enqueueBytes(queueName, value.data(), value.size());
}

void Redis::enqueueBytes(const std::string& queueName, const uint8_t* buffer, size_t bufferLen)
{
    std::lock_guard<std::mutex> lk(mx);
    lists[queueName].emplace_back(buffer, buffer + bufferLen);
    listCv.notify_all();
}

std::vector<uint8_t> Redis::popFront(const std::string& queueName, int timeoutMs)
{
    std::unique_lock<std::mutex> lk(mx);
    auto ready = [&] {
        auto it = lists.find(queueName);
        return it != lists.end() && !it->second.empty();
    };
    // timeout 0 = do not block at all (the reference switches from BLPOP to
    // LPOP: src/redis/Redis.cpp dequeueBase); negative = wait "forever"
    bool ok = timeoutMs == 0 ? ready()
                             : listCv.wait_for(lk, std::chrono::milliseconds(timeoutMs < 0 ? 3600000 : timeoutMs), ready);
    if (!ok) {
        throw RedisNoResponseException();
    }
    auto& l = lists[queueName];
    std::vector<uint8_t> v = std::move(l.front());
    l.pop_front();
    return v;
}

std::string Redis::dequeue(const std::string& queueName, int timeout)
{
    std::vector<uint8_t> v = popFront(queueName, timeout);
    return std::string(v.begin(), v.end());
}

std::vector<uint8_t> Redis::dequeueBytes(const std::string& queueName, int timeout)
{
    return popFront(queueName, timeout);
}

void Redis::dequeueBytes(const std::string& queueName, uint8_t* buffer, size_t bufferLen, int timeout)
{
    std::vector<uint8_t> v = popFront(queueName, timeout);
    if (v.size() > bufferLen) {
        throw std::runtime_error("Buffer not long enough for dequeue result (" + std::to_string(v.size()) + " > " + std::to_string(bufferLen) + ")");
    }
    memcpy(buffer, v.data(), v.size());
}

void Redis::dequeueMultiple(const std::string& queueName, uint8_t* buff, long buffLen, long nElems)
{
    // Non-destructive read of the first nElems values (LRANGE), concatenated
    std::lock_guard<std::mutex> lk(mx);
    auto it = lists.find(queueName);
    if (it == lists.end()) {
        return;
    }
    long off = 0;
    long n = 0;
    for (const auto& v : it->second) {
        if (n++ >= nElems) {
            break;
        }
        if (off + (long)v.size() > buffLen) {
            throw std::runtime_error("Buffer too small for dequeueMultiple");
        }
        memcpy(buff + off, v.data(), v.size());
        off += (long)v.size();
    }
}

void Redis::publishSchedulerResult(const std::string& key, const std::string& statusKey, const std::vector<uint8_t>& result)
{
    enqueueBytes(key, result);
    expire(key, 30);
    set(statusKey, result);
    expire(statusKey, 300);
}

} // namespace faabric::redis
