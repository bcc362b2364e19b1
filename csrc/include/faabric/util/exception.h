#pragma once

#include <stdexcept>
#include <string>

namespace faabric::util {

class FaabricException : public std::runtime_error
{
  public:
    explicit FaabricException(const std::string& message)
      : std::runtime_error(message)
    {}
};

// Thrown by an executing function when the planner told it to move elsewhere
// (reference: include/faabric/util/func.h + scheduler migration path)
class FunctionMigratedException : public FaabricException
{
  public:
    explicit FunctionMigratedException(const std::string& message)
      : FaabricException(message)
    {}
};

// Thrown when an app must be check-pointed and parked (spot eviction)
class FunctionFrozenException : public FaabricException
{
  public:
    explicit FunctionFrozenException(const std::string& message)
      : FaabricException(message)
    {}
};

} // namespace faabric::util
