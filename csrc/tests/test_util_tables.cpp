// Small util helpers, one case per scenario (reference: tests/test/util/
// test_memory.cpp "Test rounding down offsets to page size", "Test small /
// large / already aligned memory chunk", "... near page boundaries", "Test
// merging (multiple) dirty pages"; test_bytes.cpp "Test removing trailing
// zeros", "Test safe copy to small / big buffer", "Test integer encoding
// to/from bytes", "Test format byte array ...").
#include "harness.h"

#include <faabric/util/bytes.h>
#include <faabric/util/memory.h>

#include <cstring>

using namespace faabric::util;

namespace {
void addCase(const std::string& name, std::function<void()> fn)
{
    fbtest::registry().push_back({ name, "[util][table]", std::move(fn) });
}

struct ChunkRow
{
    const char* name;
    long offset;
    long length;
    long expOffsetPages;
    long expLengthPages;
    long expRemainder;
};

struct RegisterChunks
{
    RegisterChunks()
    {
        const long P = HOST_PAGE_SIZE;
        std::vector<ChunkRow> rows = {
            { "small chunk inside the first page", 10, 20, 0, 1, 10 },
            { "chunk that starts on a page boundary", 2 * P, 100, 2, 1, 0 },
            { "chunk ending exactly on a boundary", P - 50, 50, 0, 1, P - 50 },
            { "chunk straddling one boundary", P - 10, 20, 0, 2, P - 10 },
            { "chunk straddling several pages", P + 5, 3 * P, 1, 4, 5 },
            { "large offset", 100 * P + 7, 3, 100, 1, 7 },
            { "exactly one aligned page", 5 * P, P, 5, 1, 0 },
            { "one byte at the end of a page", 3 * P - 1, 1, 2, 1, P - 1 },
            { "two bytes across a boundary", 3 * P - 1, 2, 2, 2, P - 1 },
        };
        for (const auto& r : rows) {
            addCase(std::string("page aligned chunk: ") + r.name, [r, P] {
                AlignedChunk c = getPageAlignedChunk(r.offset, r.length);
                REQUIRE_EQ(c.originalOffset, r.offset);
                REQUIRE_EQ(c.originalLength, r.length);
                REQUIRE_EQ(c.nPagesOffset, r.expOffsetPages);
                REQUIRE_EQ(c.nPagesLength, r.expLengthPages);
                REQUIRE_EQ(c.nBytesOffset, r.expOffsetPages * P);
                REQUIRE_EQ(c.nBytesLength, r.expLengthPages * P);
                REQUIRE_EQ(c.offsetRemainder, r.expRemainder);
                // the aligned chunk covers the original one
                REQUIRE(c.nBytesOffset <= r.offset);
                REQUIRE(c.nBytesOffset + c.nBytesLength >= r.offset + r.length);
            });
        }
    }
} registerChunks;

struct RegisterPageMaths
{
    RegisterPageMaths()
    {
        const size_t P = (size_t)HOST_PAGE_SIZE;
        addCase("page maths: offsets round down to their page", [P] {
            REQUIRE_EQ(alignOffsetDown(0), 0u);
            REQUIRE_EQ(alignOffsetDown(1), 0u);
            REQUIRE_EQ(alignOffsetDown(P - 1), 0u);
            REQUIRE_EQ(alignOffsetDown(P), P);
            REQUIRE_EQ(alignOffsetDown(P + 1), P);
            REQUIRE_EQ(alignOffsetDown(10 * P + 123), 10 * P);
        });
        addCase("page maths: required pages round up, or down on request", [P] {
            REQUIRE_EQ(getRequiredHostPages(0), 0u);
            REQUIRE_EQ(getRequiredHostPages(1), 1u);
            REQUIRE_EQ(getRequiredHostPages(P), 1u);
            REQUIRE_EQ(getRequiredHostPages(P + 1), 2u);
            REQUIRE_EQ(getRequiredHostPages(7 * P), 7u);
            REQUIRE_EQ(getRequiredHostPagesRoundDown(P - 1), 0u);
            REQUIRE_EQ(getRequiredHostPagesRoundDown(2 * P + 5), 2u);
        });
        addCase("page maths: pointer alignment check", [P] {
            auto mem = allocatePrivateMemory(2 * P);
            REQUIRE(isPageAligned(mem.get()));
            REQUIRE(!isPageAligned(mem.get() + 1));
            REQUIRE(isPageAligned(mem.get() + P));
        });
        addCase("dirty pages: merging is an element-wise or", [] {
            std::vector<char> a = { 0, 1, 0, 0, 1 };
            std::vector<char> b = { 1, 0, 0, 1, 1 };
            mergeDirtyPages(a, b);
            REQUIRE(a == (std::vector<char>{ 1, 1, 0, 1, 1 }));
        });
        addCase("dirty pages: a longer source grows the destination", [] {
            std::vector<char> a = { 0, 1 };
            std::vector<char> b = { 1, 0, 0, 1 };
            mergeDirtyPages(a, b);
            REQUIRE(a == (std::vector<char>{ 1, 1, 0, 1 }));
            std::vector<char> empty;
            mergeDirtyPages(empty, b);
            REQUIRE(empty == b);
        });
        addCase("dirty pages: merging many vectors at once", [] {
            std::vector<char> dest = { 0, 0, 0, 0 };
            mergeManyDirtyPages(dest, { { 1, 0, 0, 0 }, { 0, 0, 1, 0 }, {}, { 0, 0, 1, 1 } });
            REQUIRE(dest == (std::vector<char>{ 1, 0, 1, 1 }));
        });
    }
} registerPageMaths;

struct RegisterBytes
{
    RegisterBytes()
    {
        addCase("bytes: trailing zeros are trimmed, inner ones kept", [] {
            std::vector<uint8_t> v = { 0, 2, 10, 0, 32, 0, 0, 0, 0 };
            trimTrailingZeros(v);
            REQUIRE(v == (std::vector<uint8_t>{ 0, 2, 10, 0, 32 }));
        });
        addCase("bytes: trimming all zeros leaves nothing", [] {
            std::vector<uint8_t> v(7, 0);
            trimTrailingZeros(v);
            REQUIRE(v.empty());
            trimTrailingZeros(v); // empty stays empty
            REQUIRE(v.empty());
        });
        addCase("bytes: safe copy into a smaller buffer truncates", [] {
            std::vector<uint8_t> data = { 0, 1, 2, 3, 4, 5 };
            uint8_t buf[3] = { 9, 9, 9 };
            int n = safeCopyToBuffer(data, buf, 3);
            REQUIRE_EQ(n, 3);
            REQUIRE(buf[0] == 0 && buf[2] == 2);
        });
        addCase("bytes: safe copy into a bigger buffer leaves the rest alone", [] {
            std::vector<uint8_t> data = { 7, 8, 9 };
            uint8_t buf[6] = { 1, 1, 1, 1, 1, 1 };
            int n = safeCopyToBuffer(data, buf, 6);
            REQUIRE_EQ(n, 3);
            REQUIRE(buf[2] == 9 && buf[3] == 1 && buf[5] == 1);
        });
        addCase("bytes: safe copy of nothing does nothing", [] {
            std::vector<uint8_t> data;
            uint8_t buf[2] = { 5, 6 };
            REQUIRE_EQ(safeCopyToBuffer(data, buf, 2), 0);
            REQUIRE(buf[0] == 5 && buf[1] == 6);
            REQUIRE_EQ(safeCopyToBuffer(nullptr, 0, buf, 2), 0);
        });
        addCase("bytes: values round-trip through byte vectors", [] {
            auto b = valueToBytes<int32_t>(-123456);
            REQUIRE_EQ(b.size(), 4u);
            REQUIRE_EQ(unalignedRead<int32_t>(b.data()), -123456);
            std::vector<uint8_t> acc;
            appendDataToBytes<uint16_t>(acc, 0xbeef);
            appendDataToBytes<double>(acc, 2.5);
            appendDataToBytes<int64_t>(acc, -7);
            REQUIRE_EQ(acc.size(), 18u);
            uint16_t a;
            double d;
            int64_t l;
            size_t off = readBytesOf(acc, 0, &a);
            off = readBytesOf(acc, off, &d);
            off = readBytesOf(acc, off, &l);
            REQUIRE_EQ(off, 18u);
            REQUIRE(a == 0xbeef && d == 2.5 && l == -7);
            REQUIRE_THROWS(readBytesOf(acc, 12, &l));
        });
        addCase("bytes: unaligned reads and writes", [] {
            uint8_t raw[16] = { 0 };
            unalignedWrite<uint32_t>(0xa1b2c3d4u, raw + 3);
            REQUIRE_EQ(unalignedRead<uint32_t>(raw + 3), 0xa1b2c3d4u);
            REQUIRE_EQ((int)raw[2], 0);
            REQUIRE_EQ((int)raw[7], 0);
            unalignedWrite<double>(-0.125, raw + 5);
            REQUIRE(unalignedRead<double>(raw + 5) == -0.125);
        });
        addCase("bytes: formatting as an int list and as hex", [] {
            std::vector<uint8_t> v = { 0, 1, 255, 16 };
            REQUIRE_EQ(formatByteArrayToIntString(v), std::string("[0, 1, 255, 16]"));
            REQUIRE_EQ(byteArrayToHexString(v.data(), (int)v.size()), std::string("0001ff10"));
            REQUIRE(hexStringToByteArray("0001ff10") == v);
            REQUIRE(hexStringToByteArray("").empty());
            REQUIRE_THROWS(hexStringToByteArray("abc")); // odd length
        });
        addCase("bytes: strings to bytes and back keep embedded zeros", [] {
            std::string s("ab\0cd", 5);
            auto b = stringToBytes(s);
            REQUIRE_EQ(b.size(), 5u);
            REQUIRE_EQ((int)b[2], 0);
            REQUIRE(bytesToString(b) == s);
        });
    }
} registerBytes;
}
