// Message <-> JSON using the schema's json names (reference: src/util/json.cpp,
// enums printed as ints, default-valued fields omitted).
#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/exception.h>

#include <string>

namespace faabric::util {

class JsonSerialisationException : public faabric::util::FaabricException
{
  public:
    explicit JsonSerialisationException(std::string message)
      : FaabricException(std::move(message))
    {}
};

template<typename M>
std::string messageToJson(const M& msg)
{
    faabric::proto::JsonWriter w;
    msg.toJson(w);
    return w.str();
}

template<typename M>
void jsonToMessage(const std::string& jsonStr, M* msg)
{
    try {
        faabric::proto::JsonValue v = faabric::proto::JsonValue::parse(jsonStr);
        msg->Clear();
        if (!msg->fromJson(v)) {
            throw JsonSerialisationException("JSON does not match message schema");
        }
    } catch (const std::runtime_error& e) {
        throw JsonSerialisationException(std::string("Bad JSON input: ") + e.what());
    }
}

}
