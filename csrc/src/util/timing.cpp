#include <faabric/util/logging.h>
#include <faabric/util/timing.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <sstream>
#include <vector>

namespace faabric::util {

struct TimerTotal
{
    long totalMicros = 0;
    long count = 0;
};

static std::mutex timerMx;
static std::map<std::string, TimerTotal> timerTotals;
static TimePoint globalTimerStart;
static bool globalTimerSet = false;

Clock& getGlobalClock()
{
    static Clock clock;
    return clock;
}

TimePoint startTimer()
{
    return getGlobalClock().now();
}

long getTimeDiffNanos(const TimePoint& begin)
{
    return getGlobalClock().timeDiffNano(getGlobalClock().now(), begin);
}

long getTimeDiffMicros(const TimePoint& begin)
{
    return getGlobalClock().timeDiffMicro(getGlobalClock().now(), begin);
}

double getTimeDiffMillis(const TimePoint& begin)
{
    return (double)getTimeDiffNanos(begin) / 1e6;
}

void logEndTimer(const std::string& label, const TimePoint& begin)
{
    long micros = getTimeDiffMicros(begin);
    {
        std::lock_guard<std::mutex> lk(timerMx);
        TimerTotal& t = timerTotals[label];
        t.totalMicros += micros;
        t.count++;
    }
    SPDLOG_TRACE("TIME = {:.3f}ms ({})", (double)micros / 1000.0, label);
}

void startGlobalTimer()
{
    std::lock_guard<std::mutex> lk(timerMx);
    globalTimerStart = getGlobalClock().now();
    globalTimerSet = true;
}

std::string getTimerTotalsString()
{
    std::vector<std::pair<std::string, TimerTotal>> rows;
    {
        std::lock_guard<std::mutex> lk(timerMx);
        rows.assign(timerTotals.begin(), timerTotals.end());
    }
    std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) {
        return a.second.totalMicros > b.second.totalMicros;
    });
    std::ostringstream os;
    for (const auto& [label, t] : rows) {
        os << label << ":" << t.totalMicros << ":" << t.count << "\n";
    }
    return os.str();
}

void printTimerTotals()
{
    std::vector<std::pair<std::string, TimerTotal>> rows;
    double totalMs = 0;
    {
        std::lock_guard<std::mutex> lk(timerMx);
        rows.assign(timerTotals.begin(), timerTotals.end());
        if (globalTimerSet) {
            totalMs = getTimeDiffMillis(globalTimerStart);
        }
    }
    std::sort(rows.begin(), rows.end(), [](const auto& a, const auto& b) {
        return a.second.totalMicros > b.second.totalMicros;
    });
    printf("---------- TIMER TOTALS ----------\n");
    printf("%-12s %-8s %s\n", "Total (ms)", "Count", "Label");
    for (const auto& [label, t] : rows) {
        printf("%-12.3f %-8ld %s\n",
               (double)t.totalMicros / 1000.0,
               t.count,
               label.c_str());
    }
    if (totalMs > 0) {
        printf("Total running time: %.3fms\n", totalMs);
    }
}

void clearTimerTotals()
{
    std::lock_guard<std::mutex> lk(timerMx);
    timerTotals.clear();
}

uint64_t timespecToNanos(struct timespec* nativeTimespec)
{
    return (uint64_t)nativeTimespec->tv_sec * 1000000000ull +
           (uint64_t)nativeTimespec->tv_nsec;
}

void nanosToTimespec(uint64_t nanos, struct timespec* nativeTimespec)
{
    nativeTimespec->tv_sec = (time_t)(nanos / 1000000000ull);
    nativeTimespec->tv_nsec = (long)(nanos % 1000000000ull);
}

} // namespace faabric::util
