// A received (or to-be-sent) transport message: 16-byte header + payload.
// Header layout matches the reference (transport/Message.h:11-21): code u8,
// size u64, sequence number i32, 3 bytes padding.
#pragma once

#include <cstdint>
#include <cstring>
#include <span>
#include <string>
#include <vector>

namespace faabric::transport {

#define NO_HEADER 0
#define HEADER_MSG_SIZE 16
#define SHUTDOWN_HEADER 220
// Sync response carrying the text of an exception thrown by the handler
#define ERROR_HEADER 221
static const std::vector<uint8_t> shutdownPayload = { 0, 0, 1, 1 };

#define NO_SEQUENCE_NUM -1

enum class MessageResponseCode
{
    SUCCESS,
    TERM,
    TIMEOUT,
    ERROR
};

class Message final
{
  public:
    Message() = default;

    // Empty message signalling an outcome (e.g. TIMEOUT)
    explicit Message(MessageResponseCode failCodeIn)
      : failCode(failCodeIn)
    {}

    Message(uint8_t codeIn, int seqIn, std::vector<uint8_t>&& payloadIn)
      : code(codeIn)
      , sequenceNum(seqIn)
      , payload(std::move(payloadIn))
    {}

    Message(uint8_t codeIn, int seqIn, const uint8_t* data, size_t size)
      : code(codeIn)
      , sequenceNum(seqIn)
      , payload(data, data + size)
    {}

    // Non-owning view of a caller's buffer: used by the in-process sync fast
    // path, where the handler runs on the caller's stack.  Anything that
    // outlives the call must ensureOwned() first.
    static Message view(uint8_t codeIn, int seqIn, const uint8_t* data, size_t size)
    {
        Message m;
        m.code = codeIn;
        m.sequenceNum = seqIn;
        m.borrowed = std::span<const uint8_t>(data, size);
        m.isView = true;
        return m;
    }

    void ensureOwned()
    {
        if (isView) {
            payload.assign(borrowed.begin(), borrowed.end());
            borrowed = {};
            isView = false;
        }
    }

    Message(Message&& other) = default;

    Message& operator=(Message&& other) = default;

    Message(const Message&) = delete;

    Message& operator=(const Message&) = delete;

    MessageResponseCode getResponseCode() const { return failCode; }

    std::vector<uint8_t> dataCopy() const
    {
        auto d = udata();
        return std::vector<uint8_t>(d.begin(), d.end());
    }

    std::span<const uint8_t> udata() const
    {
        return isView ? borrowed : std::span<const uint8_t>(payload.data(), payload.size());
    }

    std::span<const char> data() const
    {
        auto d = udata();
        return std::span<const char>((const char*)d.data(), d.size());
    }

    std::vector<uint8_t>& buffer()
    {
        ensureOwned();
        return payload;
    }

    size_t size() const { return isView ? borrowed.size() : payload.size(); }

    uint8_t getMessageCode() const { return code; }

    int getSequenceNum() const { return sequenceNum; }

    // Serialised header
    static void writeHeader(uint8_t* out, uint8_t code, uint64_t size, int32_t seq)
    {
        memset(out, 0, HEADER_MSG_SIZE);
        out[0] = code;
        memcpy(out + 1, &size, sizeof(uint64_t));
        memcpy(out + 1 + sizeof(uint64_t), &seq, sizeof(int32_t));
    }

    static void readHeader(const uint8_t* in, uint8_t& code, uint64_t& size, int32_t& seq)
    {
        code = in[0];
        memcpy(&size, in + 1, sizeof(uint64_t));
        memcpy(&seq, in + 1 + sizeof(uint64_t), sizeof(int32_t));
    }

  private:
    uint8_t code = NO_HEADER;
    int sequenceNum = NO_SEQUENCE_NUM;
    std::vector<uint8_t> payload;
    std::span<const uint8_t> borrowed;
    bool isView = false;
    MessageResponseCode failCode = MessageResponseCode::SUCCESS;
};

}
