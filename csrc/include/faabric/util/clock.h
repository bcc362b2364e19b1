#pragma once

#include <chrono>
#include <cstdint>

namespace faabric::util {

using TimePoint = std::chrono::steady_clock::time_point;

class Clock
{
  public:
    Clock() = default;

    TimePoint now() const { return std::chrono::steady_clock::now(); }

    // Wall-clock milliseconds since the Unix epoch (message timestamps)
    long epochMillis() const
    {
        return (long)std::chrono::duration_cast<std::chrono::milliseconds>(
                 std::chrono::system_clock::now().time_since_epoch())
          .count();
    }

    long epochMicros() const
    {
        return (long)std::chrono::duration_cast<std::chrono::microseconds>(
                 std::chrono::system_clock::now().time_since_epoch())
          .count();
    }

    long timeDiff(const TimePoint& t1, const TimePoint& t2) const
    {
        return (long)std::chrono::duration_cast<std::chrono::milliseconds>(t1 -
                                                                           t2)
          .count();
    }

    long timeDiffMicro(const TimePoint& t1, const TimePoint& t2) const
    {
        return (long)std::chrono::duration_cast<std::chrono::microseconds>(t1 -
                                                                           t2)
          .count();
    }

    long timeDiffNano(const TimePoint& t1, const TimePoint& t2) const
    {
        return (long)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 -
                                                                          t2)
          .count();
    }
};

Clock& getGlobalClock();

} // namespace faabric::util
