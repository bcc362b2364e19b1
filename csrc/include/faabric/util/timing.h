// Self-tracing macros, compiled in only with -DTRACE_ALL (reference:
// include/faabric/util/timing.h:6-17).  Adds a CUDA-event timer for device
// work timed on a stream.
#pragma once

#include <faabric/util/clock.h>

#include <string>

#ifdef TRACE_ALL
#define PROF_BEGIN faabric::util::startGlobalTimer();
#define PROF_START(name)                                                       \
    const faabric::util::TimePoint name = faabric::util::startTimer();
#define PROF_END(name) faabric::util::logEndTimer(#name, name);
#define PROF_SUMMARY faabric::util::printTimerTotals();
#define PROF_CLEAR faabric::util::clearTimerTotals();
#else
#define PROF_BEGIN
#define PROF_START(name)
#define PROF_END(name)
#define PROF_SUMMARY
#define PROF_CLEAR
#endif

namespace faabric::util {

TimePoint startTimer();

long getTimeDiffNanos(const TimePoint& begin);

long getTimeDiffMicros(const TimePoint& begin);

double getTimeDiffMillis(const TimePoint& begin);

void logEndTimer(const std::string& label, const TimePoint& begin);

void startGlobalTimer();

void printTimerTotals();

void clearTimerTotals();

// Returns "label:totalMicros:count" lines, sorted by total descending
std::string getTimerTotalsString();

uint64_t timespecToNanos(struct timespec* nativeTimespec);

void nanosToTimespec(uint64_t nanos, struct timespec* nativeTimespec);

} // namespace faabric::util
