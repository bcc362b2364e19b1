#pragma once

#include <faabric/transport/Message.h>
#include <faabric/transport/MessageEndpoint.h>

#include <string>

namespace faabric::transport {

// A (host, asyncPort, syncPort) client.  Endpoints are not created in mock mode
// (reference: src/transport/MessageEndpointClient.cpp:7-79).
class MessageEndpointClient
{
  public:
    MessageEndpointClient(std::string hostIn,
                          int asyncPortIn,
                          int syncPortIn,
                          int timeoutMs = DEFAULT_SOCKET_TIMEOUT_MS);

    virtual ~MessageEndpointClient() = default;

    // Serialised-message variants (any class with SerializeAsString)
    template<typename M>
    void asyncSend(int header, M* msg, int sequenceNum = NO_SEQUENCE_NUM)
    {
        std::string buffer = msg->SerializeAsString();
        asyncSend(header, (const uint8_t*)buffer.data(), buffer.size(), sequenceNum);
    }

    void asyncSend(int header,
                   const uint8_t* buffer,
                   size_t bufferSize,
                   int sequenceNum = NO_SEQUENCE_NUM);

    template<typename M, typename R>
    void syncSend(int header, M* msg, R* response)
    {
        std::string buffer = msg->SerializeAsString();
        syncSend(header, (const uint8_t*)buffer.data(), buffer.size(), response);
    }

    template<typename R>
    void syncSend(int header, const uint8_t* buffer, size_t bufferSize, R* response)
    {
        Message res = syncSendRaw(header, buffer, bufferSize);
        if (!response->ParseFromArray(res.udata().data(), (int)res.udata().size())) {
            throw std::runtime_error("Error deserialising message");
        }
    }

    Message syncSendRaw(int header, const uint8_t* buffer, size_t bufferSize);

    const std::string& getHost() const { return host; }

  protected:
    const std::string host;

  private:
    const int asyncPort;
    const int syncPort;

    AsyncSendMessageEndpoint asyncEndpoint;
    SyncSendMessageEndpoint syncEndpoint;
};

}
