// One case per case of the reference's Redis client tests, run against the
// in-process emulation of the command set faabric uses
// (reference: tests/test/redis/test_redis.cpp:15-720)
#include "fixtures.h"

#include <faabric/redis/Redis.h>
#include <faabric/util/bytes.h>

#include <thread>

using namespace tests;
using faabric::redis::Redis;

namespace {
std::vector<uint8_t> bytesOf(const std::string& s)
{
    return std::vector<uint8_t>(s.begin(), s.end());
}

void delIfEqCase(Redis& redis)
{
    redis.flushAll();
    const std::string key = "delifeq_test";
    // nothing set: nothing happens
    redis.delIfEq(key, 101);
    REQUIRE(redis.get(key).empty());
    // another value: stays
    redis.setnxex(key, 101, 60);
    redis.delIfEq(key, 102);
    REQUIRE_EQ(redis.getLong(key), 101L);
    // the actual value: gone
    redis.delIfEq(key, 101);
    REQUIRE(redis.get(key).empty());
}

void lockCase(Redis& redis)
{
    redis.flushAll();
    const std::string key = "lock_test";
    const std::string lockKey = key + "_lock";
    std::vector<uint8_t> value = { 0, 1, 2, 3 };
    redis.set(key, value);
    // releasing a lock nobody holds does nothing
    redis.releaseLock(key, 1234);
    auto lockId = redis.acquireLock(key, 10);
    REQUIRE(lockId > 0);
    REQUIRE_EQ(redis.getLong(lockKey), (long)lockId);
    REQUIRE(redis.get(key) == value);
    // nobody else gets it
    REQUIRE_EQ(redis.acquireLock(key, 10), 0u);
    REQUIRE_EQ(redis.getLong(lockKey), (long)lockId);
    // a wrong id does not release it
    redis.releaseLock(key, lockId + 1);
    REQUIRE_EQ(redis.acquireLock(key, 10), 0u);
    REQUIRE_EQ(redis.getLong(lockKey), (long)lockId);
    REQUIRE(redis.get(key) == value);
    // the right one does, and the next holder gets a new id
    redis.releaseLock(key, lockId);
    auto again = redis.acquireLock(key, 10);
    REQUIRE(again > 0);
    REQUIRE(again != lockId);
    REQUIRE_EQ(redis.getLong(lockKey), (long)again);
    REQUIRE(redis.get(key) == value);
}
}

TEST_CASE("redis case: basic operations (ping, counters, get / set, ranges, enqueue / dequeue)", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    Redis& st = Redis::getState();
    q.flushAll();
    st.ping();
    q.ping();
    // incr / decr
    REQUIRE_EQ(q.getCounter("test_counter"), 0L);
    REQUIRE_EQ(q.incr("test_counter"), 1L);
    REQUIRE_EQ(q.getCounter("test_counter"), 1L);
    REQUIRE_EQ(q.incr("test_counter"), 2L);
    REQUIRE_EQ(q.incr("test_counter"), 3L);
    REQUIRE_EQ(q.incr("test_counter"), 4L);
    REQUIRE_EQ(q.decr("test_counter"), 3L);
    REQUIRE_EQ(q.decr("test_counter"), 2L);
    REQUIRE_EQ(q.getCounter("test_counter"), 2L);
    // incrby / decrby
    REQUIRE_EQ(q.incrByLong("by", 5), 5L);
    REQUIRE_EQ(q.incrByLong("by", 10), 15L);
    REQUIRE_EQ(q.decrByLong("by", 4), 11L);
    REQUIRE_EQ(q.getCounter("by"), 11L);
    // get / set / del, into a caller's buffer too
    const auto a = bytesOf("val a"), b = bytesOf("val b");
    q.set("key a", a);
    q.set("key b", b.data(), b.size());
    REQUIRE(q.get("key a") == a);
    REQUIRE(q.get("key b") == b);
    std::vector<uint8_t> buf(a.size(), 0);
    q.get("key a", buf.data(), buf.size());
    REQUIRE(buf == a);
    q.del("key a");
    REQUIRE(q.get("key a").empty());
    REQUIRE(q.get("never set").empty());
    // longs
    q.setLong("long", 1234567890123L);
    REQUIRE_EQ(q.getLong("long"), 1234567890123L);
    REQUIRE_EQ(q.getLong("no long"), 0L);
    // set range / get range (inclusive end, like GETRANGE)
    q.set("range", { 0, 0, 0, 0, 0, 0, 0, 0 });
    uint8_t patch[3] = { 7, 8, 9 };
    q.setRange("range", 2, patch, 3);
    REQUIRE(q.get("range") == (std::vector<uint8_t>{ 0, 0, 7, 8, 9, 0, 0, 0 }));
    uint8_t part[4] = { 1, 1, 1, 1 };
    q.getRange("range", part, 4, 3, 6);
    REQUIRE(part[0] == 8 && part[1] == 9 && part[2] == 0 && part[3] == 0);
    // enqueue / dequeue keep order, per queue
    q.enqueue("my queue", "val a");
    q.enqueue("my queue", "val b");
    q.enqueue("other queue", "val c");
    REQUIRE_EQ(q.listLength("my queue"), 2L);
    REQUIRE_EQ(q.dequeue("my queue"), std::string("val a"));
    REQUIRE_EQ(q.dequeue("other queue"), std::string("val c"));
    REQUIRE_EQ(q.dequeue("my queue"), std::string("val b"));
    REQUIRE_EQ(q.listLength("my queue"), 0L);
}

TEST_CASE("redis case: strlen", "[redis][cases]")
{
    Redis& st = Redis::getState();
    st.flushAll();
    st.set("alpha", { 0, 1, 2, 3, 4, 5, 6, 7, 8 });
    st.set("beta", bytesOf("barbaz"));
    REQUIRE_EQ(st.strlen("alpha"), 9u);
    REQUIRE_EQ(st.strlen("beta"), 6u);
    REQUIRE_EQ(st.strlen("blahblah"), 0u);
}

TEST_CASE("redis case: setnxex", "[redis][cases]")
{
    Redis& st = Redis::getState();
    st.flushAll();
    REQUIRE(st.setnxex("setnxex_test", 101, 60));
    REQUIRE_EQ(st.getLong("setnxex_test"), 101L);
    REQUIRE(!st.setnxex("setnxex_test", 102, 60));
    REQUIRE_EQ(st.getLong("setnxex_test"), 101L);
    st.del("setnxex_test");
    REQUIRE(st.setnxex("setnxex_test", 102, 60));
    REQUIRE_EQ(st.getLong("setnxex_test"), 102L);
}

TEST_CASE("redis case: del if equal, state role", "[redis][cases]")
{
    delIfEqCase(Redis::getState());
}

TEST_CASE("redis case: del if equal, queue role", "[redis][cases]")
{
    delIfEqCase(Redis::getQueue());
}

TEST_CASE("redis case: acquire / release lock, state role", "[redis][cases]")
{
    lockCase(Redis::getState());
}

TEST_CASE("redis case: acquire / release lock, queue role", "[redis][cases]")
{
    lockCase(Redis::getQueue());
}

TEST_CASE("redis case: set operations on empty sets", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    REQUIRE(q.smembers("test_empty_set").empty());
}

TEST_CASE("redis case: set operations", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    // odd strings such as IPs are fine
    const std::string a = "12.45.67.89", b = "val_b", c = "192.168.3.4";
    REQUIRE_EQ(q.scard("set_a"), 0L);
    REQUIRE_EQ(q.scard("set_b"), 0L);
    REQUIRE(!q.sismember("set_a", a));
    q.sadd("set_a", a);
    q.sadd("set_a", b);
    q.sadd("set_b", c);
    REQUIRE_EQ(q.scard("set_a"), 2L);
    REQUIRE_EQ(q.scard("set_b"), 1L);
    REQUIRE(q.sismember("set_a", a));
    REQUIRE(q.sismember("set_a", b));
    REQUIRE(!q.sismember("set_a", c));
    REQUIRE(q.sismember("set_b", c));
    REQUIRE(q.smembers("set_a") == (std::set<std::string>{ a, b }));
    // adding twice changes nothing, removing takes it out
    q.sadd("set_a", a);
    REQUIRE_EQ(q.scard("set_a"), 2L);
    q.srem("set_a", a);
    REQUIRE_EQ(q.scard("set_a"), 1L);
    REQUIRE(!q.sismember("set_a", a));
    q.srem("set_a", "not there");
    REQUIRE_EQ(q.scard("set_a"), 1L);
}

TEST_CASE("redis case: random member of a set", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    REQUIRE(q.srandmember("set_foo").empty());
    q.sadd("set_foo", "val_a");
    REQUIRE_EQ(q.srandmember("set_foo"), std::string("val_a"));
    q.sadd("set_foo", "val_b");
    std::string got = q.srandmember("set_foo");
    REQUIRE((got == "val_a" || got == "val_b"));
}

TEST_CASE("redis case: set difference", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    REQUIRE(q.sdiff("set_a", "set_b").empty());
    q.sadd("set_a", "aaa");
    q.sadd("set_a", "bbb");
    q.sadd("set_a", "ccc");
    REQUIRE(q.sdiff("set_a", "set_b") == (std::set<std::string>{ "aaa", "bbb", "ccc" }));
    REQUIRE(q.sdiff("set_b", "set_a").empty());
    q.sadd("set_b", "bbb");
    q.sadd("set_b", "ddd");
    REQUIRE(q.sdiff("set_a", "set_b") == (std::set<std::string>{ "aaa", "ccc" }));
    REQUIRE(q.sdiff("set_b", "set_a") == (std::set<std::string>{ "ddd" }));
}

TEST_CASE("redis case: set intersection", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    REQUIRE(q.sinter("set_a", "set_b").empty());
    q.sadd("set_a", "aaa");
    q.sadd("set_a", "bbb");
    REQUIRE(q.sinter("set_a", "set_b").empty());
    q.sadd("set_b", "bbb");
    q.sadd("set_b", "ccc");
    REQUIRE(q.sinter("set_a", "set_b") == (std::set<std::string>{ "bbb" }));
    REQUIRE(q.sinter("set_b", "set_a") == (std::set<std::string>{ "bbb" }));
}

TEST_CASE("redis case: non-blocking dequeue on an empty queue", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    bool threwString = false, threwBytes = false;
    try {
        q.dequeue("foobar", 0);
    } catch (const faabric::redis::RedisNoResponseException&) {
        threwString = true;
    }
    try {
        q.dequeueBytes("foobar", 0);
    } catch (const faabric::redis::RedisNoResponseException&) {
        threwBytes = true;
    }
    REQUIRE(threwString);
    REQUIRE(threwBytes);
}

TEST_CASE("redis case: dequeue after enqueue, with and without a timeout", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    for (int timeout : { 0, 500 }) {
        q.flushAll();
        q.enqueue("foobar", "baz");
        REQUIRE_EQ(q.dequeue("foobar", timeout), std::string("baz"));
    }
}

TEST_CASE("redis case: enqueue after a blocking dequeue", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    std::atomic<bool> success{ false };
    std::thread waiter([&] { success = Redis::getQueue().dequeue("foobar") == "baz"; });
    std::this_thread::sleep_for(std::chrono::milliseconds(100));
    q.enqueue("foobar", "baz");
    waiter.join();
    REQUIRE(success.load());
}

TEST_CASE("redis case: enqueue and dequeue multiple", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    const std::vector<uint8_t> a = { 0, 1 }, b = { 2, 3 }, c = { 4, 5 }, d = { 6, 7 };
    for (auto* v : { &a, &b, &c, &d }) {
        q.enqueueBytes("dummyQueue", *v);
    }
    // the first two, WITHOUT taking them off the list
    std::vector<uint8_t> firstTwo(4, 9);
    q.dequeueMultiple("dummyQueue", firstTwo.data(), (long)firstTwo.size(), 2);
    REQUIRE(firstTwo == (std::vector<uint8_t>{ 0, 1, 2, 3 }));
    REQUIRE_EQ(q.listLength("dummyQueue"), 4L);
    std::vector<uint8_t> all(8, 9);
    q.dequeueMultiple("dummyQueue", all.data(), (long)all.size(), 4);
    REQUIRE(all == (std::vector<uint8_t>{ 0, 1, 2, 3, 4, 5, 6, 7 }));
}

TEST_CASE("redis case: dequeue multiple from an empty list leaves the buffer alone", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    q.del("dummyQueueEmpty");
    REQUIRE_EQ(q.listLength("dummyQueueEmpty"), 0L);
    std::vector<uint8_t> buf = { 0, 0, 0, 0 };
    q.dequeueMultiple("dummyQueueEmpty", buf.data(), (long)buf.size(), 4);
    REQUIRE(buf == (std::vector<uint8_t>{ 0, 0, 0, 0 }));
}

TEST_CASE("redis case: pipelined range sets", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    q.del("dummyPipeline");
    std::vector<uint8_t> ua = { 1, 1 }, ub = { 2, 2, 2 }, uc = { 8, 8 }, ud = { 4 };
    q.setRangePipeline("dummyPipeline", 0, ua.data(), 2);
    q.setRangePipeline("dummyPipeline", 2, ub.data(), 3);
    q.setRangePipeline("dummyPipeline", 6, uc.data(), 2);
    q.setRangePipeline("dummyPipeline", 1, ud.data(), 1);
    q.flushPipeline(4);
    std::vector<uint8_t> actual(9, 0);
    q.get("dummyPipeline", actual.data(), actual.size());
    REQUIRE(actual == (std::vector<uint8_t>{ 1, 4, 2, 2, 2, 0, 8, 8, 0 }));
}

TEST_CASE("redis case: enqueue / dequeue bytes through pointers, queues interleaved", "[redis][cases]")
{
    Redis& q = Redis::getQueue();
    q.flushAll();
    std::vector<uint8_t> a = { 0, 1, 2, 3, 4, 5 }, b = { 6, 7 }, c = { 2, 4, 6 };
    q.enqueueBytes("testQueueA", a.data(), a.size());
    q.enqueueBytes("testQueueB", b.data(), b.size());
    q.enqueueBytes("testQueueB", c.data(), c.size());
    q.enqueueBytes("testQueueA", b.data(), b.size());
    q.enqueueBytes("testQueueB", a.data(), a.size());
    auto expectNext = [&](const std::string& queue, const std::vector<uint8_t>& want) {
        std::vector<uint8_t> buf(want.size(), 0xff);
        q.dequeueBytes(queue, buf.data(), buf.size());
        REQUIRE(buf == want);
    };
    expectNext("testQueueA", a);
    expectNext("testQueueA", b);
    expectNext("testQueueB", b);
    expectNext("testQueueB", c);
    expectNext("testQueueB", a);
}
