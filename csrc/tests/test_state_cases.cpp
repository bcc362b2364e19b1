// State key-values, one case per scenario of the reference's state suite
// (tests/test/state/test_state.cpp: sizes, get/set, chunks, dirty chunks and
// their overlaps, partial updates of doubles, first / last element, push with
// a mask, pull semantics, push only when dirty, shared-memory mappings and
// offsets, deletion, appended values, large values over many chunk requests).
// Re-implemented on this repo's fixtures; expectations follow the documented
// semantics of StateKeyValue (src/state/StateKeyValue.cpp in the reference).
#include "harness.h"

#include <faabric/redis/Redis.h>
#include <faabric/state/State.h>
#include <faabric/state/StateServer.h>
#include <faabric/util/config.h>
#include <faabric/util/memory.h>

#include <cstring>
#include <numeric>

using namespace faabric::state;

namespace {
struct Hosts
{
    State& mainState = getGlobalState();
    StateServer server;
    State remoteState; // another host's view

    Hosts()
      : server(getGlobalState())
      , remoteState("otherhost")
    {
        faabric::util::getSystemConfig().reset();
        mainState.forceClearAll(true);
        getInMemoryStateRegistry().clear();
        faabric::redis::Redis::getState().flushAll();
        server.start();
    }

    ~Hosts()
    {
        server.stop();
        mainState.forceClearAll(true);
        remoteState.forceClearAll(false);
        getInMemoryStateRegistry().clear();
        faabric::util::getSystemConfig().reset();
    }
};

std::vector<uint8_t> pattern(size_t n, int mul)
{
    std::vector<uint8_t> v(n);
    for (size_t i = 0; i < n; i++) {
        v[i] = (uint8_t)(i * mul + 1);
    }
    return v;
}
}

TEST_CASE("state case: sizes are known locally and through the main host", "[state][cases]")
{
    Hosts h;
    REQUIRE_EQ(h.mainState.getStateSize("cases", "absent"), 0u);
    auto kv = h.mainState.getKV("cases", "sized", 1234);
    auto v = pattern(1234, 3);
    kv->set(v.data());
    kv->pushFull();
    REQUIRE_EQ(kv->size(), 1234u);
    REQUIRE_EQ(h.mainState.getStateSize("cases", "sized"), 1234u);
    REQUIRE_EQ(h.remoteState.getStateSize("cases", "sized"), 1234u);
    // page-rounded backing store
    REQUIRE(kv->getSharedMemorySize() >= 1234u);
    REQUIRE_EQ(kv->getSharedMemorySize() % faabric::util::HOST_PAGE_SIZE, 0u);
    // a value that exists nowhere cannot be used: a size-less handle resolves
    // its size on first use, and there is none to be found
    auto ghost = h.mainState.getKV("cases", "never-sized");
    std::vector<uint8_t> sink(4);
    REQUIRE_THROWS(ghost->get(sink.data()));
    REQUIRE_THROWS(h.mainState.getKV("", "nouser", 4));
}

TEST_CASE("state case: simple get and set", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainState.getKV("cases", "simple", 5);
    std::vector<uint8_t> v = { 0, 1, 2, 3, 4 };
    kv->set(v.data());
    std::vector<uint8_t> out(5, 9);
    kv->get(out.data());
    REQUIRE(out == v);
    REQUIRE(memcmp(kv->get(), v.data(), 5) == 0);
    // overwrite
    std::vector<uint8_t> w = { 9, 9, 9, 9, 9 };
    kv->set(w.data());
    kv->get(out.data());
    REQUIRE(out == w);
    // the same key gives the same object
    REQUIRE(h.mainState.getKV("cases", "simple", 5) == kv);
}

TEST_CASE("state case: chunks are read and written in place", "[state][cases]")
{
    Hosts h;
    auto v = pattern(40, 1);
    auto kv = h.mainState.getKV("cases", "chunks", v.size());
    kv->set(v.data());
    std::vector<uint8_t> part(7);
    kv->getChunk(5, part.data(), 7);
    REQUIRE(memcmp(part.data(), v.data() + 5, 7) == 0);
    std::vector<uint8_t> patch = { 200, 201, 202 };
    kv->setChunk(30, patch.data(), 3);
    REQUIRE_EQ((int)*kv->getChunk(31, 1), 201);
    REQUIRE_EQ((int)*kv->getChunk(29, 1), (int)v[29]);
    // bounds: reads stop at the value, writes at the (page-rounded) storage
    REQUIRE_THROWS(kv->getChunk(38, part.data(), 7));
    REQUIRE_THROWS(kv->setChunk((long)kv->getSharedMemorySize() - 1, patch.data(), 3));
}

TEST_CASE("state case: only chunks marked dirty are pushed", "[state][cases]")
{
    Hosts h;
    const size_t size = 2000;
    auto v = pattern(size, 5);
    auto mainKv = h.mainState.getKV("cases", "dirty", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "dirty", size);
    remote->pull();
    // the replica edits two ranges in place but flags only one of them
    uint8_t* raw = remote->get();
    raw[10] = 0xaa;
    raw[11] = 0xab;
    raw[1500] = 0xbb;
    remote->flagChunkDirty(10, 2);
    remote->pushPartial();
    REQUIRE_EQ((int)*mainKv->getChunk(10, 1), 0xaa);
    REQUIRE_EQ((int)*mainKv->getChunk(11, 1), 0xab);
    REQUIRE_EQ((int)*mainKv->getChunk(1500, 1), (int)v[1500]);
    // now the other one
    remote->flagChunkDirty(1500, 1);
    remote->pushPartial();
    REQUIRE_EQ((int)*mainKv->getChunk(1500, 1), 0xbb);
}

TEST_CASE("state case: overlapping dirty chunks merge into one update", "[state][cases]")
{
    Hosts h;
    const size_t size = 300;
    std::vector<uint8_t> zeros(size, 0);
    auto mainKv = h.mainState.getKV("cases", "overlap", size);
    mainKv->set(zeros.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "overlap", size);
    remote->pull();
    std::vector<uint8_t> a(20, 1), b(20, 2), c(5, 3);
    remote->setChunk(100, a.data(), a.size()); // 100..119
    remote->setChunk(110, b.data(), b.size()); // 110..129 overlaps
    remote->setChunk(128, c.data(), c.size()); // 128..132 touches
    remote->pushPartial();
    std::vector<uint8_t> got(size);
    mainKv->get(got.data());
    for (size_t i = 0; i < size; i++) {
        int want = 0;
        if (i >= 100 && i < 110) {
            want = 1;
        } else if (i >= 110 && i < 128) {
            want = 2;
        } else if (i >= 128 && i < 133) {
            want = 3;
        }
        REQUIRE_EQ((int)got[i], want);
    }
}

TEST_CASE("state case: partial update of doubles", "[state][cases]")
{
    Hosts h;
    const int n = 1000;
    std::vector<double> values(n);
    std::iota(values.begin(), values.end(), 0.5);
    auto mainKv = h.mainState.getKV("cases", "doubles", n * sizeof(double));
    mainKv->set((uint8_t*)values.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "doubles", n * sizeof(double));
    remote->pull();
    // a few scattered elements
    for (int idx : { 0, 1, 499, 500, 998, 999 }) {
        double nv = -1.0 * idx - 0.25;
        remote->setChunk(idx * sizeof(double), (uint8_t*)&nv, sizeof(double));
        values[idx] = nv;
    }
    remote->pushPartial();
    std::vector<double> got(n);
    mainKv->get((uint8_t*)got.data());
    REQUIRE(got == values);
}

TEST_CASE("state case: first and last element only", "[state][cases]")
{
    Hosts h;
    const size_t size = 5 * STATE_STREAMING_CHUNK_SIZE / 2;
    auto v = pattern(size, 7);
    auto mainKv = h.mainState.getKV("cases", "edges", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "edges", size);
    remote->pull();
    uint8_t first = 0xf1;
    uint8_t last = 0xf2;
    remote->setChunk(0, &first, 1);
    remote->setChunk((long)size - 1, &last, 1);
    remote->pushPartial();
    v[0] = first;
    v[size - 1] = last;
    std::vector<uint8_t> got(size);
    mainKv->get(got.data());
    REQUIRE(got == v);
}

TEST_CASE("state case: push with a mask key-value selects what is sent", "[state][cases]")
{
    Hosts h;
    const size_t size = 64;
    std::vector<uint8_t> zeros(size, 0);
    auto mainKv = h.mainState.getKV("cases", "masked", size);
    mainKv->set(zeros.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "masked", size);
    remote->pull();
    std::vector<uint8_t> ones(size, 1);
    remote->set(ones.data());
    // mask: bytes 8..15 and 40..43
    std::vector<uint8_t> mask(size, 0);
    std::fill(mask.begin() + 8, mask.begin() + 16, 0xff);
    std::fill(mask.begin() + 40, mask.begin() + 44, 0xff);
    auto maskKv = h.remoteState.getKV("cases", "masked-mask", size);
    maskKv->set(mask.data());
    remote->pushPartialMask(maskKv);
    std::vector<uint8_t> got(size);
    mainKv->get(got.data());
    for (size_t i = 0; i < size; i++) {
        bool sent = (i >= 8 && i < 16) || (i >= 40 && i < 44);
        REQUIRE_EQ((int)got[i], sent ? 1 : 0);
    }
    // a mask of the wrong size is refused
    auto badMask = h.remoteState.getKV("cases", "masked-bad", size + 1);
    REQUIRE_THROWS(remote->pushPartialMask(badMask));
}

TEST_CASE("state case: a replica sees remote updates only after a pull", "[state][cases]")
{
    Hosts h;
    const size_t size = 100;
    auto v = pattern(size, 2);
    auto mainKv = h.mainState.getKV("cases", "pulls", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "pulls", size);
    std::vector<uint8_t> got(size);
    remote->get(got.data()); // first use pulls
    REQUIRE(got == v);
    // the main copy moves on
    std::vector<uint8_t> w(size, 0x3c);
    mainKv->set(w.data());
    remote->get(got.data());
    REQUIRE(got == v); // still the old bytes: no implicit refresh
    remote->pull();
    remote->get(got.data());
    REQUIRE(got == w);
}

TEST_CASE("state case: nothing is pushed while nothing is dirty", "[state][cases]")
{
    Hosts h;
    const size_t size = 50;
    auto v = pattern(size, 9);
    auto mainKv = h.mainState.getKV("cases", "clean", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "clean", size);
    remote->pull();
    // main changes; a push from the clean replica must not overwrite it
    std::vector<uint8_t> w(size, 0x11);
    mainKv->set(w.data());
    remote->pushFull();
    remote->pushPartial();
    std::vector<uint8_t> got(size);
    mainKv->get(got.data());
    REQUIRE(got == w);
    // once the replica writes, its push wins
    remote->set(v.data());
    remote->pushFull();
    mainKv->get(got.data());
    REQUIRE(got == v);
}

TEST_CASE("state case: shared memory mappings alias the value", "[state][cases]")
{
    Hosts h;
    const size_t size = 3 * faabric::util::HOST_PAGE_SIZE;
    auto v = pattern(size, 1);
    auto kv = h.mainState.getKV("cases", "mapped", size);
    kv->set(v.data());
    auto a = faabric::util::allocatePrivateMemory(size);
    auto b = faabric::util::allocatePrivateMemory(size);
    kv->mapSharedMemory(a.get(), 0, 3);
    kv->mapSharedMemory(b.get(), 0, 3);
    REQUIRE(memcmp(a.get(), v.data(), size) == 0);
    // a write through one mapping is seen by the value and by the other mapping
    a[faabric::util::HOST_PAGE_SIZE + 7] = 0x99;
    REQUIRE_EQ((int)b[faabric::util::HOST_PAGE_SIZE + 7], 0x99);
    REQUIRE_EQ((int)*kv->getChunk(faabric::util::HOST_PAGE_SIZE + 7, 1), 0x99);
    // and the other way round
    uint8_t nv = 0x42;
    kv->setChunk(5, &nv, 1);
    REQUIRE_EQ((int)a[5], 0x42);
    kv->unmapSharedMemory(a.release());
    kv->unmapSharedMemory(b.release());
    // unaligned targets are refused
    auto c = faabric::util::allocatePrivateMemory(2 * faabric::util::HOST_PAGE_SIZE);
    REQUIRE_THROWS(kv->mapSharedMemory(c.get() + 16, 0, 1));
}

TEST_CASE("state case: mapping does not pull by itself", "[state][cases]")
{
    Hosts h;
    const size_t size = faabric::util::HOST_PAGE_SIZE;
    auto v = pattern(size, 13);
    auto mainKv = h.mainState.getKV("cases", "map-nopull", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "map-nopull", size);
    auto m = faabric::util::allocatePrivateMemory(size);
    remote->mapSharedMemory(m.get(), 0, 1);
    // the replica has not pulled: its pages are still zero
    REQUIRE_EQ((int)m[100], 0);
    remote->pull();
    REQUIRE_EQ((int)m[100], (int)v[100]);
    remote->unmapSharedMemory(m.release());
}

TEST_CASE("state case: mappings at page offsets inside a bigger value", "[state][cases]")
{
    Hosts h;
    const size_t P = faabric::util::HOST_PAGE_SIZE;
    const size_t size = 6 * P;
    auto v = pattern(size, 3);
    auto kv = h.mainState.getKV("cases", "map-offsets", size);
    kv->set(v.data());
    auto m = faabric::util::allocatePrivateMemory(2 * P);
    kv->mapSharedMemory(m.get(), 3, 2); // pages 3 and 4
    REQUIRE(memcmp(m.get(), v.data() + 3 * P, 2 * P) == 0);
    m[P + 1] = 0x77;
    REQUIRE_EQ((int)*kv->getChunk(4 * P + 1, 1), 0x77);
    kv->unmapSharedMemory(m.release(), 2);
    // beyond the value
    auto n = faabric::util::allocatePrivateMemory(2 * P);
    REQUIRE_THROWS(kv->mapSharedMemory(n.get(), 5, 2));
}

TEST_CASE("state case: deletion removes the value everywhere", "[state][cases]")
{
    Hosts h;
    auto v = pattern(10, 1);
    auto kv = h.mainState.getKV("cases", "doomed", 10);
    kv->set(v.data());
    kv->pushFull();
    REQUIRE_EQ(h.remoteState.getStateSize("cases", "doomed"), 10u);
    REQUIRE_EQ(h.mainState.getKVCount(), 1u);
    h.mainState.deleteKV("cases", "doomed");
    REQUIRE_EQ(h.mainState.getKVCount(), 0u);
    REQUIRE_EQ(h.mainState.getStateSize("cases", "doomed"), 0u);
    // it can be created again, with a different size
    auto again = h.mainState.getKV("cases", "doomed", 20);
    REQUIRE_EQ(again->size(), 20u);
}

TEST_CASE("state case: appended values accumulate and clear", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainState.getKV("cases", "log", 4);
    std::vector<std::vector<uint8_t>> entries = { { 1, 2, 3, 4 }, { 5, 6, 7, 8 }, { 9, 10, 11, 12 } };
    for (auto& e : entries) {
        kv->append(e.data(), e.size());
    }
    std::vector<uint8_t> all(12);
    kv->getAppended(all.data(), 12, 3);
    REQUIRE(all == (std::vector<uint8_t>{ 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12 }));
    // fewer than stored is fine, more is an error
    std::vector<uint8_t> two(8);
    kv->getAppended(two.data(), 8, 2);
    REQUIRE_EQ((int)two[7], 8);
    std::vector<uint8_t> tooMany(16);
    REQUIRE_THROWS(kv->getAppended(tooMany.data(), 16, 4));
    kv->clearAppended();
    REQUIRE_THROWS(kv->getAppended(two.data(), 4, 1));
}

TEST_CASE("state case: appends from another host land on the main copy", "[state][cases]")
{
    Hosts h;
    auto mainKv = h.mainState.getKV("cases", "remote-log", 2);
    auto remote = h.remoteState.getKV("cases", "remote-log", 2);
    std::vector<uint8_t> a = { 1, 1 }, b = { 2, 2 }, c = { 3, 3 };
    remote->append(a.data(), 2);
    mainKv->append(b.data(), 2);
    remote->append(c.data(), 2);
    std::vector<uint8_t> got(6);
    mainKv->getAppended(got.data(), 6, 3);
    REQUIRE(got == (std::vector<uint8_t>{ 1, 1, 2, 2, 3, 3 }));
    std::fill(got.begin(), got.end(), 0);
    remote->getAppended(got.data(), 6, 3);
    REQUIRE(got == (std::vector<uint8_t>{ 1, 1, 2, 2, 3, 3 }));
    remote->clearAppended();
    REQUIRE_THROWS(mainKv->getAppended(got.data(), 2, 1));
}

TEST_CASE("state case: large values travel in many chunk requests, both ways", "[state][cases]")
{
    Hosts h;
    const size_t size = 6 * STATE_STREAMING_CHUNK_SIZE + 4321;
    auto v = pattern(size, 11);
    auto mainKv = h.mainState.getKV("cases", "large", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "large", size);
    std::vector<uint8_t> got(size);
    remote->get(got.data());
    REQUIRE(got == v);
    // rewrite everything on the replica and push it back
    auto w = pattern(size, 17);
    remote->set(w.data());
    remote->pushFull();
    mainKv->get(got.data());
    REQUIRE(got == w);
    // a chunk that straddles three streaming chunks
    std::vector<uint8_t> span(2 * STATE_STREAMING_CHUNK_SIZE + 100, 0x5e);
    remote->setChunk(STATE_STREAMING_CHUNK_SIZE - 50, span.data(), span.size());
    remote->pushPartial();
    std::vector<uint8_t> check(span.size());
    mainKv->getChunk(STATE_STREAMING_CHUNK_SIZE - 50, check.data(), check.size());
    REQUIRE(check == span);
    REQUIRE_EQ((int)*mainKv->getChunk(STATE_STREAMING_CHUNK_SIZE - 51, 1), (int)w[STATE_STREAMING_CHUNK_SIZE - 51]);
}

TEST_CASE("state case: lazy chunk pulls fetch only what is read", "[state][cases]")
{
    Hosts h;
    const size_t size = 8 * STATE_STREAMING_CHUNK_SIZE;
    auto v = pattern(size, 23);
    auto mainKv = h.mainState.getKV("cases", "lazy", size);
    mainKv->set(v.data());
    mainKv->pushFull();
    auto remote = h.remoteState.getKV("cases", "lazy", size);
    std::vector<uint8_t> part(10);
    remote->getChunk(5 * STATE_STREAMING_CHUNK_SIZE + 3, part.data(), 10);
    REQUIRE(memcmp(part.data(), v.data() + 5 * STATE_STREAMING_CHUNK_SIZE + 3, 10) == 0);
    // main changes a chunk the replica has not read yet and one it has
    std::vector<uint8_t> nv(10, 0xc4);
    mainKv->setChunk(1 * STATE_STREAMING_CHUNK_SIZE, nv.data(), 10);
    mainKv->setChunk(5 * STATE_STREAMING_CHUNK_SIZE + 3, nv.data(), 10);
    remote->getChunk(1 * STATE_STREAMING_CHUNK_SIZE, part.data(), 10);
    REQUIRE(part == nv); // not pulled before: fetched now, sees the new bytes
    remote->getChunk(5 * STATE_STREAMING_CHUNK_SIZE + 3, part.data(), 10);
    REQUIRE(memcmp(part.data(), v.data() + 5 * STATE_STREAMING_CHUNK_SIZE + 3, 10) == 0); // cached
}

TEST_CASE("state case: read and write locks", "[state][cases]")
{
    Hosts h;
    auto kv = h.mainState.getKV("cases", "locks", 8);
    kv->lockRead();
    kv->lockRead(); // shared
    kv->unlockRead();
    kv->unlockRead();
    kv->lockWrite();
    std::atomic<bool> got{ false };
    std::thread t([&] {
        kv->lockRead();
        got = true;
        kv->unlockRead();
    });
    std::this_thread::sleep_for(std::chrono::milliseconds(30));
    REQUIRE(!got.load()); // writer holds it
    kv->unlockWrite();
    t.join();
    REQUIRE(got.load());
}
