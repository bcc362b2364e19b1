#include <faabric/util/delta.h>
#include <faabric/util/logging.h>

#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <sstream>
#include <stdexcept>

namespace faabric::util {

// ---------------------------------------------------------------------------
// zstd through dlopen (C API is stable: size_t-returning functions)
// ---------------------------------------------------------------------------
namespace {
struct Zstd
{
    size_t (*compressBound)(size_t) = nullptr;
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok = false;
};

const Zstd& zstd()
{
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, []() {
        void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
        if (h == nullptr) {
            return;
        }
        z.compressBound = (size_t(*)(size_t))dlsym(h, "ZSTD_compressBound");
        z.compress = (size_t(*)(void*, size_t, const void*, size_t, int))dlsym(h, "ZSTD_compress");
        z.decompress = (size_t(*)(void*, size_t, const void*, size_t))dlsym(h, "ZSTD_decompress");
        z.isError = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
        z.ok = z.compressBound && z.compress && z.decompress && z.isError;
    });
    return z;
}

template<typename T>
void put(std::vector<uint8_t>& out, T v)
{
    size_t at = out.size();
    out.resize(at + sizeof(T));
    memcpy(out.data() + at, &v, sizeof(T));
}

template<typename T>
T get(const std::vector<uint8_t>& in, size_t& pos)
{
    if (pos + sizeof(T) > in.size()) {
        throw std::runtime_error("Delta stream truncated");
    }
    T v;
    memcpy(&v, in.data() + pos, sizeof(T));
    pos += sizeof(T);
    return v;
}
}

bool deltaZstdAvailable()
{
    return zstd().ok;
}

DeltaSettings::DeltaSettings(const std::string& definition)
  : usePages(false)
  , pageSize(4096)
  , xorWithOld(false)
  , useZstd(false)
  , zstdLevel(1)
{
    std::stringstream ss(definition);
    std::string part;
    while (std::getline(ss, part, ';')) {
        if (part.empty()) {
            continue;
        }
        if (part.rfind("pages=", 0) == 0) {
            usePages = true;
            pageSize = std::stoul(part.substr(6));
            if (pageSize == 0) {
                throw std::invalid_argument("Delta page size must be > 0");
            }
        } else if (part == "xor") {
            xorWithOld = true;
        } else if (part.rfind("zstd=", 0) == 0) {
            useZstd = true;
            zstdLevel = std::stoi(part.substr(5));
        } else {
            throw std::invalid_argument("Invalid DeltaSettings configuration argument: " + part);
        }
    }
}

std::string DeltaSettings::toString() const
{
    std::stringstream ss;
    if (usePages) {
        ss << "pages=" << pageSize << ';';
    }
    if (xorWithOld) {
        ss << "xor;";
    }
    if (useZstd) {
        ss << "zstd=" << zstdLevel << ';';
    }
    return ss.str();
}

std::vector<uint8_t> serializeDelta(const DeltaSettings& cfg,
                                    const uint8_t* oldDataStart,
                                    size_t oldDataLen,
                                    const uint8_t* newDataStart,
                                    size_t newDataLen)
{
    std::vector<uint8_t> cmds;
    cmds.reserve(std::max<size_t>(newDataLen / 8, 64));
    cmds.push_back(DELTACMD_TOTAL_SIZE);
    put<uint32_t>(cmds, (uint32_t)newDataLen);

    auto emit = [&](size_t offset, size_t length) {
        if (length == 0) {
            return;
        }
        bool overlapOld = offset < oldDataLen;
        if (cfg.xorWithOld && overlapOld) {
            size_t xorLen = std::min(length, oldDataLen - offset);
            cmds.push_back(DELTACMD_DELTA_XOR);
            put<uint32_t>(cmds, (uint32_t)offset);
            put<uint32_t>(cmds, (uint32_t)xorLen);
            size_t at = cmds.size();
            cmds.resize(at + xorLen);
            for (size_t i = 0; i < xorLen; i++) {
                cmds[at + i] = newDataStart[offset + i] ^ oldDataStart[offset + i];
            }
            offset += xorLen;
            length -= xorLen;
            if (length == 0) {
                return;
            }
        }
        cmds.push_back(DELTACMD_DELTA_OVERWRITE);
        put<uint32_t>(cmds, (uint32_t)offset);
        put<uint32_t>(cmds, (uint32_t)length);
        cmds.insert(cmds.end(), newDataStart + offset, newDataStart + offset + length);
    };

    if (cfg.usePages) {
        // Merge runs of changed pages into single commands
        size_t runStart = 0;
        bool inRun = false;
        for (size_t off = 0; off < newDataLen; off += cfg.pageSize) {
            size_t len = std::min(cfg.pageSize, newDataLen - off);
            bool changed = off + len > oldDataLen ||
                           memcmp(oldDataStart + off, newDataStart + off, len) != 0;
            if (changed && !inRun) {
                inRun = true;
                runStart = off;
            } else if (!changed && inRun) {
                emit(runStart, off - runStart);
                inRun = false;
            }
        }
        if (inRun) {
            emit(runStart, newDataLen - runStart);
        }
    } else {
        emit(0, newDataLen);
    }
    return deltaFinish(cfg, std::move(cmds));
}

void deltaBegin(std::vector<uint8_t>& cmds, uint32_t totalSize)
{
    cmds.push_back(DELTACMD_TOTAL_SIZE);
    put<uint32_t>(cmds, totalSize);
}

void deltaAppendRun(std::vector<uint8_t>& cmds, bool isXor, uint32_t offset, const uint8_t* payload, uint32_t length)
{
    if (length == 0) {
        return;
    }
    cmds.push_back(isXor ? DELTACMD_DELTA_XOR : DELTACMD_DELTA_OVERWRITE);
    put<uint32_t>(cmds, offset);
    put<uint32_t>(cmds, length);
    cmds.insert(cmds.end(), payload, payload + length);
}

std::vector<uint8_t> deltaFinish(const DeltaSettings& cfg, std::vector<uint8_t>&& cmdsIn)
{
    std::vector<uint8_t> cmds = std::move(cmdsIn);
    cmds.push_back(DELTACMD_END);

    if (!cfg.useZstd || !zstd().ok) {
        return cmds;
    }
    const Zstd& z = zstd();
    size_t bound = z.compressBound(cmds.size());
    std::vector<uint8_t> out;
    out.push_back(DELTACMD_ZSTD_COMPRESSED_COMMANDS);
    size_t headerAt = out.size();
    out.resize(headerAt + 16 + bound);
    size_t n = z.compress(out.data() + headerAt + 16, bound, cmds.data(), cmds.size(), cfg.zstdLevel);
    if (z.isError(n)) {
        throw std::runtime_error("zstd compression failed");
    }
    uint64_t compLen = n;
    uint64_t rawLen = cmds.size();
    memcpy(out.data() + headerAt, &compLen, 8);
    memcpy(out.data() + headerAt + 8, &rawLen, 8);
    out.resize(headerAt + 16 + n);
    out.push_back(DELTACMD_END);
    return out;
}

void deltaForEach(const std::vector<uint8_t>& delta,
                  const std::function<void(uint32_t)>& onSize,
                  const std::function<void(bool, uint32_t, const uint8_t*, uint32_t)>& onRun)
{
    size_t pos = 0;
    while (pos < delta.size()) {
        uint8_t cmd = delta[pos++];
        switch (cmd) {
            case DELTACMD_TOTAL_SIZE: {
                onSize(get<uint32_t>(delta, pos));
                break;
            }
            case DELTACMD_ZSTD_COMPRESSED_COMMANDS: {
                uint64_t compLen = get<uint64_t>(delta, pos);
                uint64_t rawLen = get<uint64_t>(delta, pos);
                if (pos + compLen > delta.size()) {
                    throw std::runtime_error("Delta stream truncated");
                }
                const Zstd& z = zstd();
                if (!z.ok) {
                    throw std::runtime_error("zstd delta received but libzstd is unavailable");
                }
                std::vector<uint8_t> inner(rawLen);
                size_t n = z.decompress(inner.data(), rawLen, delta.data() + pos, compLen);
                if (z.isError(n) || n != rawLen) {
                    throw std::runtime_error("zstd decompression failed");
                }
                pos += compLen;
                deltaForEach(inner, onSize, onRun);
                break;
            }
            case DELTACMD_DELTA_OVERWRITE:
            case DELTACMD_DELTA_XOR: {
                uint32_t offset = get<uint32_t>(delta, pos);
                uint32_t length = get<uint32_t>(delta, pos);
                if (pos + length > delta.size()) {
                    throw std::runtime_error("Delta stream truncated");
                }
                onRun(cmd == DELTACMD_DELTA_XOR, offset, delta.data() + pos, length);
                pos += length;
                break;
            }
            case DELTACMD_END:
                return;
            default:
                throw std::runtime_error("Invalid delta command");
        }
    }
}

void applyDelta(const std::vector<uint8_t>& delta,
                std::function<void(uint32_t)> setDataSize,
                std::function<uint8_t*()> getDataPointer)
{
    deltaForEach(
      delta,
      [&](uint32_t total) { setDataSize(total); },
      [&](bool isXor, uint32_t offset, const uint8_t* payload, uint32_t length) {
          uint8_t* dst = getDataPointer() + offset;
          if (isXor) {
              for (uint32_t i = 0; i < length; i++) {
                  dst[i] ^= payload[i];
              }
          } else {
              memcpy(dst, payload, length);
          }
      });
}

} // namespace faabric::util
