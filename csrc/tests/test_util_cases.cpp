// util helpers, one case per case of the reference's suites
// (reference: tests/test/util/test_bytes.cpp, test_memory.cpp, test_queue.cpp,
// test_func.cpp)
#include "harness.h"

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/bytes.h>
#include <faabric/util/func.h>
#include <faabric/util/latch.h>
#include <faabric/util/memory.h>
#include <faabric/util/queue.h>

#include <future>
#include <thread>

using namespace faabric::util;

// ---------------------------------------------------------------------------
// bytes
// ---------------------------------------------------------------------------
TEST_CASE("bytes case: strings become bytes char by char", "[util][bytes][cases]")
{
    auto b = stringToBytes("abcde");
    REQUIRE_EQ(b.size(), 5u);
    REQUIRE('a' == (char)b[0] && 'b' == (char)b[1] && 'c' == (char)b[2] && 'd' == (char)b[3] && 'e' == (char)b[4]);
    REQUIRE_EQ(bytesToString(b), std::string("abcde"));
}

TEST_CASE("bytes case: trailing zeros are trimmed, inner ones stay", "[util][bytes][cases]")
{
    std::vector<uint8_t> v = { 0, 2, 10, 0, 32, 0, 0, 0, 0 };
    trimTrailingZeros(v);
    REQUIRE(v == (std::vector<uint8_t>{ 0, 2, 10, 0, 32 }));
}

TEST_CASE("bytes case: trimming all zeros leaves nothing", "[util][bytes][cases]")
{
    std::vector<uint8_t> v = { 0, 0, 0, 0, 0 };
    trimTrailingZeros(v);
    REQUIRE(v.empty());
}

TEST_CASE("bytes case: safe copy into a smaller buffer truncates", "[util][bytes][cases]")
{
    std::vector<uint8_t> data = { 0, 1, 2, 3, 4, 5, 6 };
    uint8_t small[3];
    REQUIRE_EQ(safeCopyToBuffer(data, small, 3), 3);
    REQUIRE(small[0] == 0 && small[1] == 1 && small[2] == 2);
}

TEST_CASE("bytes case: safe copy into a bigger buffer copies everything", "[util][bytes][cases]")
{
    std::vector<uint8_t> data = { 0, 1, 2, 3, 4, 5, 6 };
    uint8_t big[20];
    REQUIRE_EQ(safeCopyToBuffer(data, big, 20), 7);
    REQUIRE(std::vector<uint8_t>(big, big + 7) == data);
}

TEST_CASE("bytes case: safe copy of nothing touches nothing", "[util][bytes][cases]")
{
    std::vector<uint8_t> data;
    uint8_t buf[3] = { 0, 0, 0 };
    REQUIRE_EQ(safeCopyToBuffer(data, buf, 3), 0);
    REQUIRE(buf[0] == 0 && buf[1] == 0 && buf[2] == 0);
}

TEST_CASE("bytes case: safe copy keeps punctuation and survives a round trip to a C string", "[util][bytes][cases]")
{
    std::string input = "abc/def.com";
    uint8_t buf[20];
    safeCopyToBuffer(stringToBytes(input), buf, 20);
    buf[11] = '\0';
    REQUIRE_EQ(std::string(reinterpret_cast<char*>(buf)), input);
    REQUIRE('/' == (char)buf[3] && '.' == (char)buf[7] && 'm' == (char)buf[10]);
}

TEST_CASE("bytes case: integers of every width are appended and read back in order", "[util][bytes][cases]")
{
    std::vector<uint8_t> buf;
    uint8_t v1 = 2, r1;
    uint16_t v2 = 0xABCD, r2;
    uint32_t v4 = 0xBEEF1337, r4;
    uint64_t v8 = 0xABEE2929BEE51234ull, r8;
    appendBytesOf(buf, v1);
    REQUIRE_EQ(buf.size(), 1u);
    appendBytesOf(buf, v2);
    REQUIRE_EQ(buf.size(), 3u);
    appendBytesOf(buf, v4);
    REQUIRE_EQ(buf.size(), 7u);
    appendBytesOf(buf, v8);
    REQUIRE_EQ(buf.size(), 15u);
    size_t off = 0;
    off = readBytesOf(buf, off, &r1);
    REQUIRE(r1 == v1 && off == 1);
    off = readBytesOf(buf, off, &r2);
    REQUIRE(r2 == v2 && off == 3);
    off = readBytesOf(buf, off, &r4);
    REQUIRE(r4 == v4 && off == 7);
    off = readBytesOf(buf, off, &r8);
    REQUIRE(r8 == v8 && off == 15);
    bool ranged = false;
    try {
        readBytesOf(buf, off, &r1);
    } catch (const std::range_error&) {
        ranged = true;
    }
    REQUIRE(ranged);
}

TEST_CASE("bytes case: byte arrays print as lists of ints", "[util][bytes][cases]")
{
    REQUIRE_EQ(formatByteArrayToIntString({}), std::string("[]"));
    REQUIRE_EQ(formatByteArrayToIntString({ 0, 1, 2, 3, 4, 5, 6, 7 }), std::string("[0, 1, 2, 3, 4, 5, 6, 7]"));
    REQUIRE_EQ(formatByteArrayToIntString({ 23, 9, 100 }), std::string("[23, 9, 100]"));
}

TEST_CASE("bytes case: hex strings and byte arrays convert both ways", "[util][bytes][cases]")
{
    std::vector<std::pair<std::vector<uint8_t>, std::string>> table = {
        { { 0, 1, 2, 3, 4, 5, 6, 7 }, "0001020304050607" },
        { { 'F', 'O', 'O', 12, 2, 3, 4, 5, 6, 7 }, "464f4f0c020304050607" },
        { { 'F', '*', 12, '_', ')' }, "462a0c5f29" },
    };
    for (auto& [bytes, hex] : table) {
        REQUIRE_EQ(byteArrayToHexString(bytes.data(), (int)bytes.size()), hex);
        REQUIRE(hexStringToByteArray(hex) == bytes);
    }
}

TEST_CASE("bytes case: ints print as zero-padded hex of their width", "[util][bytes][cases]")
{
    REQUIRE_EQ(intToHexString<uint8_t>(0), std::string("00"));
    REQUIRE_EQ(intToHexString<uint8_t>(255), std::string("ff"));
    REQUIRE_EQ(intToHexString<uint16_t>(255), std::string("00ff"));
    REQUIRE_EQ(intToHexString<uint32_t>(0xBEEF1337), std::string("beef1337"));
    REQUIRE_EQ(intToHexString<uint64_t>(1), std::string("0000000000000001"));
    REQUIRE_EQ(intToHexString<uint64_t>(0xABEE2929BEE51234ull), std::string("abee2929bee51234"));
}

TEST_CASE("bytes case: four bytes make an int", "[util][bytes][cases]")
{
    int v = 0x11223344;
    std::vector<uint8_t> b = valueToBytes<int>(v);
    REQUIRE_EQ(b.size(), sizeof(int));
    REQUIRE_EQ(bytesToInt(b), v);
    REQUIRE_EQ(bytesToInt(valueToBytes<int>(-7)), -7);
}

// ---------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------
TEST_CASE("memory case: offsets round down to page boundaries", "[util][memory][cases]")
{
    REQUIRE_EQ(alignOffsetDown(0), 0u);
    REQUIRE_EQ(alignOffsetDown(1), 0u);
    REQUIRE_EQ(alignOffsetDown(HOST_PAGE_SIZE - 1), 0u);
    REQUIRE_EQ(alignOffsetDown(HOST_PAGE_SIZE), (size_t)HOST_PAGE_SIZE);
    REQUIRE_EQ(alignOffsetDown(2 * HOST_PAGE_SIZE + 33), (size_t)(2 * HOST_PAGE_SIZE));
    REQUIRE_EQ(getRequiredHostPages(1), 1u);
    REQUIRE_EQ(getRequiredHostPages(HOST_PAGE_SIZE), 1u);
    REQUIRE_EQ(getRequiredHostPages(HOST_PAGE_SIZE + 1), 2u);
    REQUIRE_EQ(getRequiredHostPagesRoundDown(2 * HOST_PAGE_SIZE - 1), 1u);
}

TEST_CASE("memory case: a small chunk at offset zero covers one page", "[util][memory][cases]")
{
    AlignedChunk c = getPageAlignedChunk(0, 10);
    REQUIRE(c.originalOffset == 0 && c.originalLength == 10);
    REQUIRE(c.nBytesOffset == 0 && c.nBytesLength == HOST_PAGE_SIZE);
    REQUIRE(c.nPagesOffset == 0 && c.nPagesLength == 1 && c.offsetRemainder == 0);
}

TEST_CASE("memory case: a chunk straddling a page boundary covers both pages", "[util][memory][cases]")
{
    AlignedChunk c = getPageAlignedChunk(2 * HOST_PAGE_SIZE - 1, 3);
    REQUIRE(c.nPagesOffset == 1 && c.nPagesLength == 2);
    REQUIRE(c.nBytesOffset == HOST_PAGE_SIZE && c.nBytesLength == 2 * HOST_PAGE_SIZE);
    REQUIRE_EQ(c.offsetRemainder, HOST_PAGE_SIZE - 1);
}

TEST_CASE("memory case: a large chunk at a large unaligned offset", "[util][memory][cases]")
{
    AlignedChunk c = getPageAlignedChunk(2 * HOST_PAGE_SIZE + 33, 5 * HOST_PAGE_SIZE + 123);
    REQUIRE(c.nPagesOffset == 2 && c.nPagesLength == 6);
    REQUIRE(c.nBytesOffset == 2 * HOST_PAGE_SIZE && c.nBytesLength == 6 * HOST_PAGE_SIZE);
    REQUIRE_EQ(c.offsetRemainder, 33L);
}

TEST_CASE("memory case: an already aligned chunk is left as it is", "[util][memory][cases]")
{
    AlignedChunk c = getPageAlignedChunk(10 * HOST_PAGE_SIZE, 5 * HOST_PAGE_SIZE);
    REQUIRE(c.nPagesOffset == 10 && c.nPagesLength == 5);
    REQUIRE(c.nBytesOffset == 10 * HOST_PAGE_SIZE && c.nBytesLength == 5 * HOST_PAGE_SIZE);
    REQUIRE_EQ(c.offsetRemainder, 0L);
}

TEST_CASE("memory case: virtual memory is reserved first and usable once claimed", "[util][memory][cases]")
{
    size_t size = 10 * HOST_PAGE_SIZE;
    MemoryRegion region = allocateVirtualMemory(size);
    REQUIRE(region != nullptr);
    REQUIRE(isPageAligned(region.get()));
    // claim the first half and use it
    claimVirtualMemory({ region.get(), size / 2 });
    region.get()[0] = 5;
    region.get()[size / 2 - 1] = 6;
    REQUIRE(region.get()[0] == 5 && region.get()[size / 2 - 1] == 6);
    // then the rest
    claimVirtualMemory({ region.get() + size / 2, size / 2 });
    region.get()[size - 1] = 7;
    REQUIRE_EQ((int)region.get()[size - 1], 7);
}

TEST_CASE("memory case: private mappings of an fd are copy-on-write, shared ones write through", "[util][memory][cases]")
{
    size_t size = 4 * HOST_PAGE_SIZE;
    std::vector<uint8_t> data(size, 3);
    int fd = createFd(size, "memcase");
    writeToFd(fd, 0, { data.data(), data.size() });
    MemoryRegion priv = allocatePrivateMemory(size);
    MemoryRegion sharedA = allocateSharedMemory(size);
    MemoryRegion sharedB = allocateSharedMemory(size);
    mapMemoryPrivate({ priv.get(), size }, fd);
    mapMemoryShared({ sharedA.get(), size }, fd);
    mapMemoryShared({ sharedB.get(), size }, fd);
    REQUIRE(std::vector<uint8_t>(priv.get(), priv.get() + size) == data);
    REQUIRE(std::vector<uint8_t>(sharedA.get(), sharedA.get() + size) == data);
    // a private write stays private
    priv.get()[10] = 9;
    REQUIRE_EQ((int)sharedA.get()[10], 3);
    // a shared write is seen by the other shared mapping, not by the private one's touched page
    sharedA.get()[HOST_PAGE_SIZE + 5] = 8;
    REQUIRE_EQ((int)sharedB.get()[HOST_PAGE_SIZE + 5], 8);
    REQUIRE_EQ((int)priv.get()[10], 9);
    ::close(fd);
}

TEST_CASE("memory case: mapping from a zero or negative fd is refused", "[util][memory][cases]")
{
    size_t size = 10 * HOST_PAGE_SIZE;
    MemoryRegion mem = allocatePrivateMemory(size);
    REQUIRE_THROWS(mapMemoryPrivate({ mem.get(), size }, 0));
    REQUIRE_THROWS(mapMemoryPrivate({ mem.get(), size }, -2));
    REQUIRE_THROWS(mapMemoryShared({ mem.get(), size }, -2));
}

TEST_CASE("memory case: remapping a private mapping drops the changes made through it", "[util][memory][cases]")
{
    size_t size = 10 * HOST_PAGE_SIZE;
    std::vector<uint8_t> expected(size, 3);
    int fd = createFd(size, "foobar");
    writeToFd(fd, 0, { expected.data(), expected.size() });
    MemoryRegion mem = allocatePrivateMemory(size);
    mapMemoryPrivate({ mem.get(), size }, fd);
    REQUIRE(std::vector<uint8_t>(mem.get(), mem.get() + size) == expected);
    std::vector<uint8_t> update(100, 4);
    size_t off = HOST_PAGE_SIZE + 10;
    memcpy(mem.get() + off, update.data(), update.size());
    REQUIRE_EQ((int)mem.get()[off + 5], 4);
    mapMemoryPrivate({ mem.get(), size }, fd);
    REQUIRE_EQ((int)mem.get()[off + 5], 3);
    REQUIRE(std::vector<uint8_t>(mem.get(), mem.get() + size) == expected);
    ::close(fd);
}

TEST_CASE("memory case: merging dirty page flags of equal and unequal lengths", "[util][memory][cases]")
{
    {
        std::vector<char> src = { 0, 1, 0, 1, 1, 0 }, dst = { 0, 1, 1, 0, 1, 0 };
        mergeDirtyPages(dst, src);
        REQUIRE(dst == (std::vector<char>{ 0, 1, 1, 1, 1, 0 }));
    }
    {
        // a longer source grows the destination
        std::vector<char> src = { 0, 1, 0, 1, 1, 0, 1 }, dst = { 0, 1, 1 };
        mergeDirtyPages(dst, src);
        REQUIRE(dst == (std::vector<char>{ 0, 1, 1, 1, 1, 0, 1 }));
    }
    {
        // a shorter one only touches what it covers
        std::vector<char> src = { 1, 0 }, dst = { 0, 0, 1, 0 };
        mergeDirtyPages(dst, src);
        REQUIRE(dst == (std::vector<char>{ 1, 0, 1, 0 }));
    }
    {
        std::vector<char> dst;
        mergeDirtyPages(dst, {});
        REQUIRE(dst.empty());
    }
}

TEST_CASE("memory case: merging several sets of dirty page flags at once", "[util][memory][cases]")
{
    std::vector<char> dst = { 0, 0, 0, 0 };
    std::vector<std::vector<char>> many = { { 1, 0, 0, 0 }, { 0, 0, 1 }, { 0, 0, 0, 0, 1 } };
    mergeManyDirtyPages(dst, many);
    REQUIRE(dst == (std::vector<char>{ 1, 0, 1, 0, 1 }));
    mergeManyDirtyPages(dst, {});
    REQUIRE(dst == (std::vector<char>{ 1, 0, 1, 0, 1 }));
}

// ---------------------------------------------------------------------------
// queues
// ---------------------------------------------------------------------------
TEST_CASE("queue case: fifo order, size, peek and present-only dequeue", "[util][queue][cases]")
{
    Queue<int> q;
    q.enqueue(1);
    q.enqueue(2);
    q.enqueue(3);
    REQUIRE_EQ(q.size(), 3L);
    REQUIRE_EQ(*q.peek(), 1);
    REQUIRE_EQ(q.dequeue(), 1);
    REQUIRE_EQ(q.dequeue(), 2);
    int out = -1;
    q.dequeueIfPresent(&out);
    REQUIRE_EQ(out, 3);
    out = -1;
    q.dequeueIfPresent(&out); // empty: left alone
    REQUIRE_EQ(out, -1);
    bool timedOut = false;
    try {
        q.dequeue(20);
    } catch (const QueueTimeoutException&) {
        timedOut = true;
    }
    REQUIRE(timedOut);
}

TEST_CASE("queue case: drain empties the queue", "[util][queue][cases]")
{
    Queue<int> q;
    for (int i = 0; i < 5; i++) {
        q.enqueue(i);
    }
    REQUIRE_EQ(q.size(), 5L);
    q.drain();
    REQUIRE_EQ(q.size(), 0L);
    q.drain(); // again: still fine
}

TEST_CASE("queue case: waiting for an empty queue to drain returns at once", "[util][queue][cases]")
{
    Queue<int> q;
    q.waitToDrain(100);
    REQUIRE_EQ(q.size(), 0L);
}

TEST_CASE("queue case: waiting for a queue to drain blocks until a consumer took everything", "[util][queue][cases]")
{
    Queue<int> q;
    const int n = 10;
    for (int i = 0; i < n; i++) {
        q.enqueue(i);
    }
    std::vector<int> got;
    std::thread consumer([&] {
        for (int i = 0; i < n; i++) {
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            got.push_back(q.dequeue());
        }
    });
    q.waitToDrain(5000);
    REQUIRE_EQ(q.size(), 0L);
    consumer.join();
    REQUIRE_EQ((int)got.size(), n);
    // waiting on a queue nobody drains times out
    q.enqueue(1);
    bool timedOut = false;
    try {
        q.waitToDrain(20);
    } catch (const QueueTimeoutException&) {
        timedOut = true;
    }
    REQUIRE(timedOut);
}

TEST_CASE("queue case: move-only elements travel through both queue types", "[util][queue][cases]")
{
    auto exercise = [](auto& q) {
        std::promise<int32_t> a, b;
        auto fa = a.get_future(), fb = b.get_future();
        q.enqueue(std::move(a));
        q.enqueue(std::move(b));
        std::thread ta([&q] { q.dequeue().set_value(1); });
        ta.join();
        std::thread tb([&q] { q.dequeue().set_value(2); });
        tb.join();
        REQUIRE_EQ(fa.get(), 1);
        REQUIRE_EQ(fb.get(), 2);
    };
    Queue<std::promise<int32_t>> plain;
    exercise(plain);
    FixedCapacityQueue<std::promise<int32_t>> fixed(4);
    exercise(fixed);
}

TEST_CASE("queue case: dequeue timeouts must be positive", "[util][queue][cases]")
{
    Queue<int> q;
    q.enqueue(10);
    REQUIRE_THROWS(q.dequeue(0));
    REQUIRE_THROWS(q.dequeue(-1));
    FixedCapacityQueue<int> f(2);
    f.enqueue(10);
    REQUIRE_THROWS(f.dequeue(0));
    REQUIRE_THROWS(f.dequeue(-1));
}

TEST_CASE("queue case: a full fixed-capacity queue makes producers wait", "[util][queue][cases]")
{
    FixedCapacityQueue<int> q(2);
    q.enqueue(1);
    q.enqueue(2);
    bool timedOut = false;
    try {
        q.enqueue(100, 50);
    } catch (const QueueTimeoutException&) {
        timedOut = true;
    }
    REQUIRE(timedOut);
}

TEST_CASE("queue case: a consumer makes room in a full fixed-capacity queue", "[util][queue][cases]")
{
    FixedCapacityQueue<int> q(2);
    auto latch = Latch::create(2);
    std::thread consumer([&] {
        latch->wait();
        q.dequeue();
    });
    q.enqueue(1);
    q.enqueue(2);
    latch->wait();
    q.enqueue(3, 2000);
    consumer.join();
    REQUIRE_EQ(q.dequeue(), 2);
    REQUIRE_EQ(q.dequeue(), 3);
}

TEST_CASE("queue case: fixed-capacity queue under stress, fast and slow sides swapped", "[util][queue][cases]")
{
    for (bool slowConsumer : { false, true }) {
        FixedCapacityQueue<int> q(8);
        const int n = 2000;
        long long sum = 0;
        std::thread producer([&] {
            for (int i = 0; i < n; i++) {
                if (!slowConsumer && i % 200 == 0) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(1));
                }
                q.enqueue(i);
            }
        });
        std::thread consumer([&] {
            int last = -1;
            for (int i = 0; i < n; i++) {
                if (slowConsumer && i % 200 == 0) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(1));
                }
                int v = q.dequeue();
                if (v != last + 1) {
                    sum = -1000000000;
                }
                last = v;
                sum += v;
            }
        });
        producer.join();
        consumer.join();
        REQUIRE_EQ(sum, (long long)n * (n - 1) / 2);
    }
}

// ---------------------------------------------------------------------------
// func
// ---------------------------------------------------------------------------
TEST_CASE("func case: the message factory fills ids, keys, host and timestamp", "[util][func][cases]")
{
    faabric::Message msg = messageFactory("demo", "echo");
    REQUIRE_EQ(msg.user(), std::string("demo"));
    REQUIRE_EQ(msg.function(), std::string("echo"));
    REQUIRE(msg.id() > 0);
    REQUIRE(msg.appid() > 0);
    REQUIRE(msg.starttimestamp() > 0);
    REQUIRE_EQ(msg.resultkey(), resultKeyFromMessageId(msg.id()));
    REQUIRE_EQ(msg.statuskey(), statusKeyFromMessageId(msg.id()));
    REQUIRE(!msg.mainhost().empty());
    auto shared = messageFactoryShared("demo", "echo");
    REQUIRE_EQ(shared->user(), std::string("demo"));
    REQUIRE(shared->id() > 0);
    REQUIRE(shared->id() != msg.id());
}

TEST_CASE("func case: setting the id of a message without one", "[util][func][cases]")
{
    faabric::Message a, b;
    REQUIRE_EQ(a.id(), 0);
    unsigned int idA = setMessageId(a), idB = setMessageId(b);
    REQUIRE(idA > 0 && idB > 0 && idA != idB);
    REQUIRE_EQ((unsigned int)a.id(), idA);
    REQUIRE_EQ(a.resultkey(), resultKeyFromMessageId(idA));
    REQUIRE_EQ(a.statuskey(), statusKeyFromMessageId(idA));
}

TEST_CASE("func case: a message that has an id keeps it and still gets its keys", "[util][func][cases]")
{
    faabric::Message m;
    m.set_id(1234);
    REQUIRE_EQ(setMessageId(m), 1234u);
    REQUIRE_EQ(m.id(), 1234);
    REQUIRE_EQ(m.resultkey(), resultKeyFromMessageId(1234));
    REQUIRE_EQ(m.statuskey(), statusKeyFromMessageId(1234));
    // calling it again changes nothing
    REQUIRE_EQ(setMessageId(m), 1234u);
}

TEST_CASE("func case: the asynchronous response of a call is its id", "[util][func][cases]")
{
    faabric::Message m = messageFactory("foo", "bar");
    REQUIRE_EQ(buildAsyncResponse(m), std::to_string(m.id()));
    REQUIRE_EQ(funcToString(m, false), std::string("foo/bar"));
    REQUIRE_EQ(funcToString(m, true), "foo/bar:" + std::to_string(m.id()));
}
