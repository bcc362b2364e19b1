// (reference: include/faabric/util/chaining.h)
#pragma once

#include <faabric/proto/faabric.pb.h>
#include <faabric/util/exception.h>

#include <string>

namespace faabric::util {

// Thrown by a function whose chained call came back with an error
class ChainedCallFailedException : public faabric::util::FaabricException
{
  public:
    explicit ChainedCallFailedException(std::string message)
      : FaabricException(std::move(message))
    {}
};

}
