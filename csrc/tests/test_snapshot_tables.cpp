// Byte-exact tables for the snapshot diff / merge machinery: one registered
// case per scenario (the reference spells these out as Catch2 SECTIONs of
// tests/test/util/test_snapshot.cpp: "Test diffing byte array regions", "Test
// snapshot merge regions", "Detailed test snapshot merge regions with ints",
// "Test edge-cases of snapshot merge regions", "Test fine-grained byte-wise
// diffs", "Test filling gaps in regions", "Test invalid snapshot merges", ...).
// The scenarios and expectations are re-derived from the documented semantics,
// the data tables are this repo's own.
#include "harness.h"

#include <faabric/util/bytes.h>
#include <faabric/util/config.h>
#include <faabric/util/memory.h>
#include <faabric/util/snapshot.h>

#include <cstring>
#include <map>

using namespace faabric::util;

namespace {
// ---- registration of generated cases ----
void addCase(const std::string& name, const char* tags, std::function<void()> fn)
{
    fbtest::registry().push_back({ name, tags, std::move(fn) });
}

struct ConfGuard
{
    explicit ConfGuard(const std::string& mode) { getSystemConfig().diffingMode = mode; }
    ~ConfGuard() { getSystemConfig().reset(); }
};

// ===========================================================================
// diffArrayRegions
// ===========================================================================
struct ArrayDiffRow
{
    const char* name;
    std::vector<uint8_t> a;
    std::vector<uint8_t> b;
    uint64_t start;
    uint64_t end; // 0 => b.size()
    std::vector<std::pair<uint64_t, uint64_t>> expected;
};

std::vector<uint8_t> ramp(size_t n, int mul = 1)
{
    std::vector<uint8_t> v(n);
    for (size_t i = 0; i < n; i++) {
        v[i] = (uint8_t)(i * mul);
    }
    return v;
}

std::vector<uint8_t> edited(std::vector<uint8_t> v, std::initializer_list<std::pair<size_t, size_t>> runs)
{
    for (auto [off, len] : runs) {
        for (size_t i = off; i < off + len; i++) {
            v[i] ^= 0x5a;
        }
    }
    return v;
}

std::vector<ArrayDiffRow> arrayDiffRows()
{
    auto base = ramp(1000, 3);
    return {
        { "equal arrays give no runs", { 0, 1, 2, 3 }, { 0, 1, 2, 3 }, 0, 0, {} },
        { "empty arrays give no runs", {}, {}, 0, 0, {} },
        { "two runs in the middle", { 0, 0, 2, 2, 3, 3, 4, 4, 5, 5 }, { 0, 1, 1, 2, 3, 6, 6, 6, 5, 5 }, 0, 0, { { 1, 2 }, { 5, 3 } } },
        { "window clips runs on both sides",
          { 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0 },
          { 1, 1, 1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0 },
          3,
          10,
          { { 3, 2 }, { 6, 1 }, { 8, 2 } } },
        { "single differing byte", { 0, 1, 2, 3, 4 }, { 0, 1, 3, 3, 4 }, 0, 0, { { 2, 1 } } },
        { "run at the very start", { 0, 1, 2, 3, 4, 5, 6 }, { 1, 2, 3, 3, 3, 4, 6 }, 0, 0, { { 0, 3 }, { 4, 2 } } },
        { "run reaching the very end", { 0, 1, 2, 3, 4, 5, 6 }, { 0, 1, 1, 3, 3, 4, 5 }, 0, 0, { { 2, 1 }, { 4, 3 } } },
        { "everything differs", { 1, 1, 1, 1 }, { 2, 2, 2, 2 }, 0, 0, { { 0, 4 } } },
        { "window that sees no difference", { 9, 0, 0, 0, 9 }, { 1, 0, 0, 0, 1 }, 1, 4, {} },
        { "empty window", { 1, 2, 3 }, { 3, 2, 1 }, 2, 2, {} },
        { "run across a 128-byte chunk boundary", base, edited(base, { { 120, 20 } }), 0, 0, { { 120, 20 } } },
        { "runs in non-adjacent chunks only", base, edited(base, { { 5, 1 }, { 700, 3 } }), 0, 0, { { 5, 1 }, { 700, 3 } } },
        { "one-byte gap keeps runs apart", base, edited(base, { { 300, 4 }, { 305, 2 } }), 0, 0, { { 300, 4 }, { 305, 2 } } },
        { "last byte of the last (partial) chunk", base, edited(base, { { 999, 1 } }), 0, 0, { { 999, 1 } } },
        { "window starting inside a run", base, edited(base, { { 250, 20 } }), 260, 400, { { 260, 10 } } },
        { "window ending inside a run", base, edited(base, { { 250, 20 } }), 0, 255, { { 250, 5 } } },
        { "whole chunk differs, neighbours equal", base, edited(base, { { 128, 128 } }), 0, 0, { { 128, 128 } } },
    };
}

struct RegisterArrayDiffs
{
    RegisterArrayDiffs()
    {
        for (const auto& row : arrayDiffRows()) {
            addCase(std::string("diffArrayRegions: ") + row.name, "[util][snapshot][table]", [row] {
                std::vector<std::pair<uint64_t, uint64_t>> got;
                uint64_t end = row.end == 0 ? row.b.size() : row.end;
                diffArrayRegions(got, row.start, end, row.a, row.b);
                REQUIRE_EQ(got.size(), row.expected.size());
                for (size_t i = 0; i < got.size(); i++) {
                    REQUIRE_EQ(got[i].first, row.expected[i].first);
                    REQUIRE_EQ(got[i].second, row.expected[i].second);
                }
            });
        }
    }
} registerArrayDiffs;

// ===========================================================================
// typed merge regions: (type, op) matrix with two writers and a main copy that
// moved in the meantime
// ===========================================================================
template<typename T>
void typedMergeCase(SnapshotDataType dt, SnapshotMergeOperation op, T original, T writerA, T writerB, T expected)
{
    ConfGuard conf("bytewise");
    const size_t size = 2 * HOST_PAGE_SIZE;
    const uint64_t off = HOST_PAGE_SIZE + 24;
    std::vector<uint8_t> init(size, 0);
    memcpy(init.data() + off, &original, sizeof(T));
    auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()));
    snap->addMergeRegion(off, sizeof(T), dt, op);
    snap->fillGapsWithBytewiseRegions();
    std::vector<char> dirty(2, 1);
    std::vector<std::vector<uint8_t>> keep;
    for (T v : { writerA, writerB }) {
        keep.emplace_back(init);
        auto& mem = keep.back();
        memcpy(mem.data() + off, &v, sizeof(T));
        auto diffs = snap->diffWithDirtyRegions(mem, dirty);
        if (v == original) {
            REQUIRE(diffs.empty());
            continue;
        }
        REQUIRE_EQ(diffs.size(), 1u);
        REQUIRE_EQ(diffs[0].getOffset(), off);
        REQUIRE((int)diffs[0].getOperation() == (int)op);
        REQUIRE((int)diffs[0].getDataType() == (int)dt);
        REQUIRE_EQ(diffs[0].getData().size(), sizeof(T));
        snap->queueDiffs(diffs);
    }
    snap->writeQueuedDiffs();
    T got;
    memcpy(&got, snap->getDataPtr(off), sizeof(T));
    if constexpr (std::is_floating_point_v<T>) {
        REQUIRE_NEAR((double)got, (double)expected, 1e-4 * std::max(1.0, std::fabs((double)expected)));
    } else {
        REQUIRE_EQ(got, expected);
    }
    // nothing else in the image moved
    auto all = snap->getDataCopy();
    for (size_t i = 0; i < size; i++) {
        if (i < off || i >= off + sizeof(T)) {
            REQUIRE_EQ((int)all[i], 0);
        }
    }
}

struct RegisterTypedMerges
{
    RegisterTypedMerges()
    {
        const char* tags = "[util][snapshot][table]";
        using O = SnapshotMergeOperation;
        using D = SnapshotDataType;
        // ints
        addCase("typed merge: int sum of two writers", tags, [] { typedMergeCase<int32_t>(D::Int, O::Sum, 100, 130, 95, 125); });
        addCase("typed merge: int sum with an unchanged writer", tags, [] { typedMergeCase<int32_t>(D::Int, O::Sum, 7, 7, 20, 20); });
        addCase("typed merge: int subtract", tags, [] { typedMergeCase<int32_t>(D::Int, O::Subtract, 100, 90, 60, 50); });
        addCase("typed merge: int product", tags, [] { typedMergeCase<int32_t>(D::Int, O::Product, 3, 6, 12, 24); });
        addCase("typed merge: int max keeps the largest", tags, [] { typedMergeCase<int32_t>(D::Int, O::Max, 10, 40, 25, 40); });
        addCase("typed merge: int max below the original changes nothing", tags, [] { typedMergeCase<int32_t>(D::Int, O::Max, 10, 4, 9, 10); });
        addCase("typed merge: int min keeps the smallest", tags, [] { typedMergeCase<int32_t>(D::Int, O::Min, 10, 4, 7, 4); });
        addCase("typed merge: int sum of negative deltas", tags, [] { typedMergeCase<int32_t>(D::Int, O::Sum, -5, -25, -6, -26); });
        // longs
        addCase("typed merge: long sum beyond 32 bits", tags, [] {
            typedMergeCase<int64_t>(D::Long, O::Sum, (int64_t)1 << 40, ((int64_t)1 << 40) + 5, ((int64_t)1 << 40) + ((int64_t)1 << 33), ((int64_t)1 << 40) + 5 + ((int64_t)1 << 33));
        });
        addCase("typed merge: long subtract", tags, [] { typedMergeCase<int64_t>(D::Long, O::Subtract, 1000, 400, 900, 300); });
        addCase("typed merge: long product", tags, [] { typedMergeCase<int64_t>(D::Long, O::Product, 10, 30, 20, 60); });
        addCase("typed merge: long max", tags, [] { typedMergeCase<int64_t>(D::Long, O::Max, -3, -1, -2, -1); });
        addCase("typed merge: long min", tags, [] { typedMergeCase<int64_t>(D::Long, O::Min, 50, 60, 20, 20); });
        // floats
        addCase("typed merge: float sum", tags, [] { typedMergeCase<float>(D::Float, O::Sum, 1.5f, 2.0f, 4.5f, 5.0f); });
        addCase("typed merge: float subtract", tags, [] { typedMergeCase<float>(D::Float, O::Subtract, 10.0f, 7.5f, 9.0f, 6.5f); });
        addCase("typed merge: float product", tags, [] { typedMergeCase<float>(D::Float, O::Product, 2.0f, 3.0f, 5.0f, 7.5f); });
        addCase("typed merge: float max", tags, [] { typedMergeCase<float>(D::Float, O::Max, 0.25f, 0.5f, 0.125f, 0.5f); });
        addCase("typed merge: float min", tags, [] { typedMergeCase<float>(D::Float, O::Min, 0.25f, 0.5f, 0.125f, 0.125f); });
        // doubles
        addCase("typed merge: double sum", tags, [] { typedMergeCase<double>(D::Double, O::Sum, 1e10, 1e10 + 1.25, 1e10 - 0.5, 1e10 + 0.75); });
        addCase("typed merge: double subtract", tags, [] { typedMergeCase<double>(D::Double, O::Subtract, 5.0, 4.0, 2.0, 1.0); });
        addCase("typed merge: double product", tags, [] { typedMergeCase<double>(D::Double, O::Product, 4.0, 2.0, 8.0, 4.0); });
        addCase("typed merge: double max", tags, [] { typedMergeCase<double>(D::Double, O::Max, -1.0, -0.5, -2.0, -0.5); });
        addCase("typed merge: double min", tags, [] { typedMergeCase<double>(D::Double, O::Min, -1.0, -0.5, -2.0, -2.0); });
    }
} registerTypedMerges;

// ===========================================================================
// gap filling
// ===========================================================================
struct GapRow
{
    const char* name;
    const char* mode;
    size_t snapSize;
    std::vector<SnapshotMergeRegion> in;
    std::vector<SnapshotMergeRegion> expected;
};

struct RegisterGaps
{
    RegisterGaps()
    {
        using O = SnapshotMergeOperation;
        using D = SnapshotDataType;
        const size_t S = 5 * HOST_PAGE_SIZE;
        std::vector<GapRow> rows = {
            { "no regions: one region to the end", "bytewise", S, {}, { { 0, 0, D::Raw, O::Bytewise } } },
            { "no regions, xor mode", "xor", S, {}, { { 0, 0, D::Raw, O::XOR } } },
            { "one region in the middle", "bytewise", S, { { 100, 4, D::Int, O::Sum } },
              { { 0, 100, D::Raw, O::Bytewise }, { 100, 4, D::Int, O::Sum }, { 104, 0, D::Raw, O::Bytewise } } },
            { "region at offset zero", "bytewise", S, { { 0, 8, D::Long, O::Max } },
              { { 0, 8, D::Long, O::Max }, { 8, 0, D::Raw, O::Bytewise } } },
            { "region ending exactly at the end", "bytewise", S, { { S - 8, 8, D::Double, O::Min } },
              { { 0, S - 8, D::Raw, O::Bytewise }, { S - 8, 8, D::Double, O::Min } } },
            { "adjacent regions leave no gap between them", "bytewise", S, { { 64, 4, D::Int, O::Sum }, { 68, 4, D::Int, O::Product } },
              { { 0, 64, D::Raw, O::Bytewise }, { 64, 4, D::Int, O::Sum }, { 68, 4, D::Int, O::Product }, { 72, 0, D::Raw, O::Bytewise } } },
            { "unsorted input is sorted first", "bytewise", S, { { 4096, 4, D::Int, O::Sum }, { 16, 4, D::Float, O::Sum } },
              { { 0, 16, D::Raw, O::Bytewise }, { 16, 4, D::Float, O::Sum }, { 20, 4076, D::Raw, O::Bytewise }, { 4096, 4, D::Int, O::Sum }, { 4100, 0, D::Raw, O::Bytewise } } },
            { "zero-length region swallows the tail", "bytewise", S, { { 200, 0, D::Raw, O::Ignore } },
              { { 0, 200, D::Raw, O::Bytewise }, { 200, 0, D::Raw, O::Ignore } } },
            { "xor mode fills gaps with xor regions", "xor", S, { { 128, 8, D::Long, O::Sum } },
              { { 0, 128, D::Raw, O::XOR }, { 128, 8, D::Long, O::Sum }, { 136, 0, D::Raw, O::XOR } } },
            { "ignore region in the middle keeps its neighbours bytewise", "bytewise", S, { { 1000, 500, D::Raw, O::Ignore } },
              { { 0, 1000, D::Raw, O::Bytewise }, { 1000, 500, D::Raw, O::Ignore }, { 1500, 0, D::Raw, O::Bytewise } } },
        };
        for (const auto& row : rows) {
            addCase(std::string("merge region gaps: ") + row.name, "[util][snapshot][table]", [row] {
                ConfGuard conf(row.mode);
                SnapshotData snap(row.snapSize);
                for (const auto& r : row.in) {
                    snap.addMergeRegion(r.offset, r.length, r.dataType, r.operation);
                }
                snap.fillGapsWithBytewiseRegions();
                auto got = snap.getMergeRegions();
                REQUIRE_EQ(got.size(), row.expected.size());
                for (size_t i = 0; i < got.size(); i++) {
                    REQUIRE(got[i] == row.expected[i]);
                }
                // filling twice changes nothing
                snap.fillGapsWithBytewiseRegions();
                REQUIRE_EQ(snap.getMergeRegions().size(), row.expected.size());
            });
        }
    }
} registerGaps;

// ===========================================================================
// fine-grained bytewise / xor diffs through dirty pages
// ===========================================================================
struct ByteRow
{
    const char* name;
    const char* mode;
    std::vector<std::pair<size_t, size_t>> edits;         // (offset, length) of changed bytes
    std::vector<char> dirtyPages;                          // per page; empty => all dirty
    std::vector<std::pair<uint64_t, uint64_t>> expected;   // (offset, length) of diffs
};

struct RegisterByteDiffs
{
    RegisterByteDiffs()
    {
        const size_t P = HOST_PAGE_SIZE;
        std::vector<ByteRow> rows = {
            { "no edits, no diffs", "bytewise", {}, {}, {} },
            { "single byte", "bytewise", { { 10, 1 } }, {}, { { 10, 1 } } },
            { "two runs in one page", "bytewise", { { 10, 3 }, { 200, 7 } }, {}, { { 10, 3 }, { 200, 7 } } },
            { "run crossing a page boundary is split per page", "bytewise", { { P - 4, 8 } }, {}, { { P - 4, 4 }, { P, 4 } } },
            { "edits in a page not flagged dirty are not seen", "bytewise", { { 5, 2 }, { P + 5, 2 } }, { 1, 0, 1, 1 }, { { 5, 2 } } },
            { "first and last byte of the image", "bytewise", { { 0, 1 }, { 4 * P - 1, 1 } }, {}, { { 0, 1 }, { 4 * P - 1, 1 } } },
            { "a whole page", "bytewise", { { 2 * P, P } }, {}, { { 2 * P, P } } },
            { "xor mode ships whole dirty pages", "xor", { { 3, 1 }, { 3 * P + 9, 2 } }, { 1, 0, 0, 1 }, { { 0, P }, { 3 * P, P } } },
            { "xor mode, nothing dirty", "xor", { { 3, 1 } }, { 0, 0, 0, 0 }, {} },
        };
        for (const auto& row : rows) {
            addCase(std::string("byte diffs: ") + row.name, "[util][snapshot][table]", [row, P] {
                ConfGuard conf(row.mode);
                const size_t size = 4 * P;
                std::vector<uint8_t> init = ramp(size, 7);
                auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()));
                snap->fillGapsWithBytewiseRegions();
                std::vector<uint8_t> mem = init;
                for (auto [off, len] : row.edits) {
                    for (size_t i = off; i < off + len; i++) {
                        mem[i] ^= 0xff;
                    }
                }
                std::vector<uint8_t> updated = mem; // xor diffs overwrite the memory they are computed from
                std::vector<char> dirty = row.dirtyPages.empty() ? std::vector<char>(4, 1) : row.dirtyPages;
                auto diffs = snap->diffWithDirtyRegions(mem, dirty);
                std::map<uint64_t, uint64_t> got;
                for (auto& d : diffs) {
                    got[d.getOffset()] = d.getData().size();
                }
                REQUIRE_EQ(got.size(), row.expected.size());
                for (auto [off, len] : row.expected) {
                    REQUIRE(got.count(off) == 1);
                    REQUIRE_EQ(got[off], len);
                }
                // applying the diffs to the main copy reproduces the edits the tracker saw
                snap->queueDiffs(diffs);
                snap->writeQueuedDiffs();
                auto image = snap->getDataCopy();
                for (size_t page = 0; page < 4; page++) {
                    bool seen = dirty[page] != 0;
                    const uint8_t* want = seen ? updated.data() : init.data();
                    REQUIRE(memcmp(image.data() + page * P, want + page * P, P) == 0);
                }
            });
        }
    }
} registerByteDiffs;
}

// ===========================================================================
// hand-written cases
// ===========================================================================
TEST_CASE("snapshot: constructors, sizes and bounds", "[util][snapshot]")
{
    SnapshotData empty(3 * HOST_PAGE_SIZE);
    REQUIRE_EQ(empty.getSize(), 3 * HOST_PAGE_SIZE);
    REQUIRE_EQ(empty.getMaxSize(), 3 * HOST_PAGE_SIZE);
    SnapshotData growable(HOST_PAGE_SIZE, 8 * HOST_PAGE_SIZE);
    REQUIRE_EQ(growable.getSize(), HOST_PAGE_SIZE);
    REQUIRE_EQ(growable.getMaxSize(), 8 * HOST_PAGE_SIZE);
    std::vector<uint8_t> bytes = { 1, 2, 3, 4, 5 };
    SnapshotData fromData(bytes);
    REQUIRE_EQ(fromData.getSize(), 5u);
    REQUIRE_EQ((int)fromData.getDataCopy()[4], 5);
    SnapshotData fromDataMax(bytes, 2 * HOST_PAGE_SIZE);
    REQUIRE_EQ(fromDataMax.getMaxSize(), 2 * HOST_PAGE_SIZE);
    // reads out of bounds are refused, writes may extend up to the maximum
    REQUIRE_THROWS(fromData.getDataCopy(3, 10));
    REQUIRE_THROWS(empty.getDataPtr(3 * HOST_PAGE_SIZE + 1));
    std::vector<uint8_t> more(100, 9);
    growable.copyInData(more, 2 * HOST_PAGE_SIZE);
    REQUIRE_EQ(growable.getSize(), 2 * HOST_PAGE_SIZE + 100);
    REQUIRE_THROWS(growable.copyInData(more, 8 * HOST_PAGE_SIZE - 50));
}

TEST_CASE("snapshot: growing keeps old contents and tracks the new bytes", "[util][snapshot]")
{
    std::vector<uint8_t> first(HOST_PAGE_SIZE, 0x11);
    SnapshotData snap(first, 6 * HOST_PAGE_SIZE);
    snap.clearTrackedChanges();
    std::vector<uint8_t> extra(300, 0x22);
    snap.copyInData(extra, 3 * HOST_PAGE_SIZE);
    REQUIRE_EQ(snap.getSize(), 3 * HOST_PAGE_SIZE + 300);
    REQUIRE_EQ((int)*snap.getDataPtr(10), 0x11);
    REQUIRE_EQ((int)*snap.getDataPtr(3 * HOST_PAGE_SIZE + 299), 0x22);
    // the gap that appeared is zero
    REQUIRE_EQ((int)*snap.getDataPtr(2 * HOST_PAGE_SIZE), 0);
    auto tracked = snap.getTrackedChanges();
    REQUIRE(!tracked.empty());
    bool covers = false;
    for (auto& d : tracked) {
        covers = covers || (d.getOffset() <= 3 * HOST_PAGE_SIZE && d.getOffset() + d.getData().size() >= 3 * HOST_PAGE_SIZE + 300);
    }
    REQUIRE(covers);
}

TEST_CASE("snapshot: mapping, editing the mapping and remapping restores the image", "[util][snapshot]")
{
    const size_t size = 3 * HOST_PAGE_SIZE;
    std::vector<uint8_t> init = ramp(size, 5);
    auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()));
    auto mem = allocatePrivateMemory(size);
    snap->mapToMemory({ mem.get(), size });
    REQUIRE(memcmp(mem.get(), init.data(), size) == 0);
    // the mapping is private: edits do not reach the image
    mem[100] = 0xee;
    mem[2 * HOST_PAGE_SIZE + 1] = 0xdd;
    REQUIRE_EQ((int)*snap->getDataPtr(100), (int)init[100]);
    // a change of the image is visible after a remap, the private edits are gone
    std::vector<uint8_t> patch = { 7, 7, 7 };
    snap->copyInData(patch, HOST_PAGE_SIZE);
    snap->mapToMemory({ mem.get(), size });
    REQUIRE_EQ((int)mem[100], (int)init[100]);
    REQUIRE_EQ((int)mem[HOST_PAGE_SIZE + 2], 7);
    // partial mapping of the first page only
    auto small = allocatePrivateMemory(HOST_PAGE_SIZE);
    snap->mapToMemory({ small.get(), HOST_PAGE_SIZE });
    REQUIRE(memcmp(small.get(), snap->getDataPtr(), HOST_PAGE_SIZE) == 0);
    // targets larger than the image or not page aligned are refused
    auto big = allocatePrivateMemory(4 * HOST_PAGE_SIZE);
    REQUIRE_THROWS(snap->mapToMemory({ big.get(), 4 * HOST_PAGE_SIZE }));
    REQUIRE_THROWS(snap->mapToMemory({ mem.get() + 8, HOST_PAGE_SIZE }));
}

TEST_CASE("snapshot: clearing merge regions and region ordering / equality", "[util][snapshot]")
{
    SnapshotData snap(2 * HOST_PAGE_SIZE);
    snap.addMergeRegion(500, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    snap.addMergeRegion(20, 8, SnapshotDataType::Long, SnapshotMergeOperation::Min);
    REQUIRE_EQ(snap.getMergeRegions().size(), 2u);
    snap.clearMergeRegions();
    REQUIRE(snap.getMergeRegions().empty());

    SnapshotMergeRegion a(10, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    SnapshotMergeRegion b(10, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    SnapshotMergeRegion c(10, 4, SnapshotDataType::Int, SnapshotMergeOperation::Max);
    SnapshotMergeRegion d(12, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    SnapshotMergeRegion e(10, 8, SnapshotDataType::Long, SnapshotMergeOperation::Sum);
    REQUIRE(a == b);
    REQUIRE(!(a == c));
    REQUIRE(!(a == d));
    REQUIRE(!(a == e));
    std::vector<SnapshotMergeRegion> regions = { d, e, a };
    std::sort(regions.begin(), regions.end());
    REQUIRE_EQ(regions.front().offset, 10u);
    REQUIRE_EQ(regions.back().offset, 12u);
}

TEST_CASE("snapshot: merge regions that cannot be applied are rejected", "[util][snapshot]")
{
    ConfGuard conf("bytewise");
    const size_t size = 2 * HOST_PAGE_SIZE;
    std::vector<uint8_t> init(size, 0);
    std::vector<char> dirty(2, 1);
    auto tryRegion = [&](SnapshotDataType dt, SnapshotMergeOperation op, size_t len) {
        auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()));
        snap->addMergeRegion(64, len, dt, op);
        std::vector<uint8_t> mem = init;
        mem[64] = 1;
        mem[65] = 2;
        return snap->diffWithDirtyRegions(mem, dirty);
    };
    // raw data only knows bytewise / xor / ignore
    REQUIRE_THROWS(tryRegion(SnapshotDataType::Raw, SnapshotMergeOperation::Sum, 16));
    REQUIRE_THROWS(tryRegion(SnapshotDataType::Raw, SnapshotMergeOperation::Max, 16));
    // booleans have no arithmetic
    REQUIRE_THROWS(tryRegion(SnapshotDataType::Bool, SnapshotMergeOperation::Product, 1));
    // the supported combinations go through
    REQUIRE_EQ(tryRegion(SnapshotDataType::Int, SnapshotMergeOperation::Sum, 4).size(), 1u);
    REQUIRE_EQ(tryRegion(SnapshotDataType::Raw, SnapshotMergeOperation::Bytewise, 16).size(), 1u);
    REQUIRE(tryRegion(SnapshotDataType::Raw, SnapshotMergeOperation::Ignore, 16).empty());
}

TEST_CASE("snapshot: memory that grew past the image is diffed as an extension", "[util][snapshot]")
{
    ConfGuard conf("bytewise");
    const size_t snapSize = 2 * HOST_PAGE_SIZE;
    const size_t memSize = 4 * HOST_PAGE_SIZE;
    std::vector<uint8_t> init(snapSize, 3);
    auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()), 8 * HOST_PAGE_SIZE);
    snap->fillGapsWithBytewiseRegions();
    std::vector<uint8_t> mem(memSize, 3);
    mem[10] = 4;                                  // inside the image
    memset(mem.data() + snapSize, 0, memSize - snapSize);
    mem[3 * HOST_PAGE_SIZE + 5] = 9;              // in the extension
    std::vector<char> dirty(4, 1);
    auto diffs = snap->diffWithDirtyRegions(mem, dirty);
    bool sawInside = false;
    bool sawExtension = false;
    for (auto& d : diffs) {
        sawInside = sawInside || (d.getOffset() == 10 && d.getData().size() == 1);
        sawExtension = sawExtension || (d.getOffset() >= snapSize);
    }
    REQUIRE(sawInside);
    REQUIRE(sawExtension);
    snap->queueDiffs(diffs);
    snap->writeQueuedDiffs();
    REQUIRE_EQ(snap->getSize(), memSize);
    REQUIRE_EQ((int)*snap->getDataPtr(3 * HOST_PAGE_SIZE + 5), 9);
    REQUIRE_EQ((int)*snap->getDataPtr(10), 4);
}

TEST_CASE("snapshot: applicable and non-applicable regions side by side", "[util][snapshot]")
{
    ConfGuard conf("bytewise");
    const size_t size = 3 * HOST_PAGE_SIZE;
    std::vector<uint8_t> init(size, 0);
    int v = 50;
    memcpy(init.data() + 2 * HOST_PAGE_SIZE + 16, &v, 4);
    auto snap = std::make_shared<SnapshotData>(std::span<const uint8_t>(init.data(), init.size()));
    // one region in a clean page, one in a dirty page, one beyond the memory
    snap->addMergeRegion(32, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    snap->addMergeRegion(2 * HOST_PAGE_SIZE + 16, 4, SnapshotDataType::Int, SnapshotMergeOperation::Sum);
    std::vector<uint8_t> mem = init;
    int nv = 80;
    memcpy(mem.data() + 2 * HOST_PAGE_SIZE + 16, &nv, 4);
    int other = 5;
    memcpy(mem.data() + 32, &other, 4); // changed, but its page is not flagged
    std::vector<char> dirty = { 0, 0, 1 };
    auto diffs = snap->diffWithDirtyRegions(mem, dirty);
    REQUIRE_EQ(diffs.size(), 1u);
    REQUIRE_EQ(diffs[0].getOffset(), 2 * HOST_PAGE_SIZE + 16);
    REQUIRE_EQ(unalignedRead<int>(diffs[0].getData().data()), 30);
}
