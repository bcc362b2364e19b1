#pragma once

#include <faabric/redis/Redis.h>
#include <faabric/state/StateKeyValue.h>

namespace faabric::state {

// STATE_MODE=redis: the value lives in the Redis-compatible store
// (reference: src/state/RedisStateKeyValue.cpp:15-129)
class RedisStateKeyValue final : public StateKeyValue
{
  public:
    RedisStateKeyValue(const std::string& userIn, const std::string& keyIn, size_t sizeIn);

    RedisStateKeyValue(const std::string& userIn, const std::string& keyIn);

    static size_t getStateSizeFromRemote(const std::string& userIn, const std::string& keyIn);

    static void deleteFromRemote(const std::string& userIn, const std::string& keyIn);

    static void clearAll(bool global);

  private:
    const std::string joinedKey;

    size_t sizeFromRemote() override;

    void pullFromRemote() override;

    void pullChunkFromRemote(long offset, size_t length) override;

    void pushToRemote() override;

    void pushPartialToRemote(const std::vector<StateChunk>& dirtyChunks) override;

    void appendToRemote(const uint8_t* data, size_t length) override;

    void pullAppendedFromRemote(uint8_t* data, size_t length, long nValues) override;

    void clearAppendedFromRemote() override;
};

}
