// Host snapshots: a memfd-backed memory image with typed merge regions, diffing
// against dirty pages and diff application.  Behavioural contract follows the
// reference (include/faabric/util/snapshot.h:21-346, src/util/snapshot.cpp);
// offsets are 64-bit here (the reference's uint32 offsets cap images at 4 GiB).
// The device-resident counterpart is faabric::snapshot::DeviceSnapshot.
#pragma once

#include <faabric/util/bytes.h>
#include <faabric/util/locks.h>
#include <faabric/util/memory.h>

#include <cstdint>
#include <map>
#include <memory>
#include <deque>
#include <span>
#include <string>
#include <vector>

namespace faabric::util {

// Granularity of the array comparison inside a dirty page
#define ARRAY_COMP_CHUNK_SIZE 128

// Numeric values are part of the application ABI
enum SnapshotDataType
{
    Raw,
    Bool,
    Int,
    Long,
    Float,
    Double
};

enum SnapshotMergeOperation
{
    Bytewise,
    Sum,
    Product,
    Subtract,
    Max,
    Min,
    Ignore,
    XOR
};

std::string snapshotDataTypeStr(SnapshotDataType dt);

std::string snapshotMergeOpStr(SnapshotMergeOperation op);

// A modification to a snapshot.  Non-owning: `data` points into the memory
// the diff was computed from (or into a kept-alive transport message).
class SnapshotDiff
{
  public:
    SnapshotDiff() = default;

    SnapshotDiff(SnapshotDataType dataTypeIn,
                 SnapshotMergeOperation operationIn,
                 uint64_t offsetIn,
                 std::span<const uint8_t> dataIn);

    SnapshotDataType getDataType() const { return dataType; }

    SnapshotMergeOperation getOperation() const { return operation; }

    uint64_t getOffset() const { return offset; }

    std::span<const uint8_t> getData() const { return data; }

    std::vector<uint8_t> getDataCopy() const;

  private:
    SnapshotDataType dataType = SnapshotDataType::Raw;
    SnapshotMergeOperation operation = SnapshotMergeOperation::Bytewise;
    uint64_t offset = 0;
    std::span<const uint8_t> data;
};

class SnapshotMergeRegion
{
  public:
    uint64_t offset = 0;
    uint64_t length = 0; // 0 => until the end of the original data
    SnapshotDataType dataType = SnapshotDataType::Raw;
    SnapshotMergeOperation operation = SnapshotMergeOperation::Bytewise;

    SnapshotMergeRegion() = default;

    SnapshotMergeRegion(uint64_t offsetIn,
                        uint64_t lengthIn,
                        SnapshotDataType dataTypeIn,
                        SnapshotMergeOperation operationIn);

    // Appends the diffs this region produces.  NB: XOR and the typed operations
    // overwrite `updatedData` with the value to transmit (zero-copy diffs).
    void addDiffs(std::vector<SnapshotDiff>& diffs,
                  std::span<const uint8_t> originalData,
                  std::span<uint8_t> updatedData,
                  const std::vector<char>& dirtyRegions);

    bool operator<(const SnapshotMergeRegion& other) const
    {
        return offset < other.offset;
    }

    bool operator==(const SnapshotMergeRegion& other) const
    {
        return offset == other.offset && length == other.length &&
               dataType == other.dataType && operation == other.operation;
    }
};

// Value sent for a typed region (Sum: new-old, Subtract: old-new, Product:
// new/old, Max/Min: new).  Writes it over `updated`; false if unchanged.
template<typename T>
bool calculateDiffValue(const uint8_t* original,
                        uint8_t* updated,
                        SnapshotMergeOperation operation);

// Merges a received typed value into the main copy
template<typename T>
T applyDiffValue(const uint8_t* original,
                 const uint8_t* diff,
                 SnapshotMergeOperation operation);

// Byte-exact runs of difference between a and b over [startOffset, endOffset):
// 128-byte chunks are skipped by memcmp, inside a differing chunk a run ends at
// the first equal byte.  Appends (offset, length) pairs.
void diffArrayRegions(std::vector<std::pair<uint64_t, uint64_t>>& diffs,
                      uint64_t startOffset,
                      uint64_t endOffset,
                      std::span<const uint8_t> a,
                      std::span<const uint8_t> b);

class SnapshotData
{
  public:
    SnapshotData() = default;

    explicit SnapshotData(size_t sizeIn);

    SnapshotData(size_t sizeIn, size_t maxSizeIn);

    explicit SnapshotData(std::span<const uint8_t> dataIn);

    SnapshotData(std::span<const uint8_t> dataIn, size_t maxSizeIn);

    SnapshotData(const SnapshotData&) = delete;

    SnapshotData& operator=(const SnapshotData&) = delete;

    ~SnapshotData();

    void copyInData(std::span<const uint8_t> buffer, uint64_t offset = 0);

    const uint8_t* getDataPtr(uint64_t offset = 0);

    std::vector<uint8_t> getDataCopy();

    std::vector<uint8_t> getDataCopy(uint64_t offset, size_t dataSize);

    // Private copy-on-write mapping of the image onto page-aligned `target`
    void mapToMemory(std::span<uint8_t> target);

    void addMergeRegion(uint64_t offset,
                        size_t length,
                        SnapshotDataType dataType,
                        SnapshotMergeOperation operation);

    // Gap filler type follows the DIFFING_MODE config (bytewise | xor)
    void fillGapsWithBytewiseRegions();

    void clearMergeRegions();

    std::vector<SnapshotMergeRegion> getMergeRegions();

    size_t getQueuedDiffsCount();

    void queueDiffs(const std::vector<SnapshotDiff>& diffs);

    // Applies and clears the queue; returns how many were written
    int writeQueuedDiffs();

    void applyDiffs(const std::vector<SnapshotDiff>& diffs);

    void applyDiff(const SnapshotDiff& diff);

    size_t getSize() const { return size; }

    size_t getMaxSize() const { return maxSize; }

    // Every write since the last clear as Raw/Bytewise diffs into the image
    std::vector<SnapshotDiff> getTrackedChanges();

    void clearTrackedChanges();

    std::vector<SnapshotDiff> diffWithDirtyRegions(
      std::span<uint8_t> updated,
      const std::vector<char>& dirtyRegions);

  private:
    size_t size = 0;
    size_t maxSize = 0;
    int fd = -1;

    std::shared_mutex snapMx;

    MemoryRegion data = nullptr;

    std::vector<SnapshotDiff> queuedDiffs;
    std::deque<std::vector<uint8_t>> queuedDiffData;

    // offset -> end (exclusive)
    std::vector<std::pair<uint64_t, uint64_t>> trackedChanges;

    std::vector<SnapshotMergeRegion> mergeRegions;

    void init(size_t initialSize, size_t maxSizeIn);

    uint8_t* validatedOffsetPtr(uint64_t offset);

    void checkWriteExtension(std::span<const uint8_t> buffer, uint64_t offset);

    void writeData(std::span<const uint8_t> buffer, uint64_t offset = 0);

    void xorData(std::span<const uint8_t> buffer, uint64_t offset = 0);

    void applyDiffLocked(const SnapshotDiff& diff);
};

} // namespace faabric::util
