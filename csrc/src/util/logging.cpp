#include <faabric/util/config.h>
#include <faabric/util/logging.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <ctime>
#include <mutex>
#include <sys/syscall.h>
#include <unistd.h>

namespace faabric::util {

static std::atomic<int> currentLevel{ (int)LogLevel::info };
static std::mutex logMx;
static FILE* logSink = nullptr;

static LogLevel parseLevel(const std::string& name)
{
    if (name == "trace") {
        return LogLevel::trace;
    }
    if (name == "debug") {
        return LogLevel::debug;
    }
    if (name == "info") {
        return LogLevel::info;
    }
    if (name == "warn" || name == "warning") {
        return LogLevel::warn;
    }
    if (name == "error" || name == "err") {
        return LogLevel::err;
    }
    if (name == "critical") {
        return LogLevel::critical;
    }
    if (name == "off") {
        return LogLevel::off;
    }
    return LogLevel::info;
}

void initLogging()
{
    // The config constructor itself logs nothing, so this cannot recurse
    SystemConfig& conf = getSystemConfig();
    currentLevel.store((int)parseLevel(conf.logLevel));
    std::lock_guard<std::mutex> lk(logMx);
    // Re-initialising switches sinks: back to stderr, or on to another file
    if (logSink != nullptr) {
        fclose(logSink);
        logSink = nullptr;
    }
    if (conf.logFile != "off" && conf.logFile != "on" && !conf.logFile.empty()) {
        FILE* f = fopen(conf.logFile.c_str(), "a");
        if (f != nullptr) {
            logSink = f;
        }
    }
}

LogLevel getLogLevel()
{
    static std::once_flag once;
    std::call_once(once, []() { initLogging(); });
    return (LogLevel)currentLevel.load(std::memory_order_relaxed);
}

void setLogLevel(LogLevel level)
{
    getLogLevel();
    currentLevel.store((int)level);
}

void setLogLevel(const std::string& name)
{
    setLogLevel(parseLevel(name));
}

void logLine(LogLevel level, const std::string& msg)
{
    static const char* tags = "TDIWEC";
    auto now = std::chrono::system_clock::now();
    time_t t = std::chrono::system_clock::to_time_t(now);
    int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(
                     now.time_since_epoch())
                     .count() %
                   1000);
    struct tm tmv;
    localtime_r(&t, &tmv);
    static thread_local long tid = syscall(SYS_gettid);
    char head[64];
    snprintf(head,
             sizeof(head),
             "[%02d:%02d:%02d.%03d] [%ld] [%c] ",
             tmv.tm_hour,
             tmv.tm_min,
             tmv.tm_sec,
             ms,
             tid,
             tags[(int)level]);
    std::lock_guard<std::mutex> lk(logMx);
    FILE* out = logSink != nullptr ? logSink : stderr;
    fputs(head, out);
    fputs(msg.c_str(), out);
    fputc('\n', out);
    if ((int)level >= (int)LogLevel::warn) {
        fflush(out);
    }
}

} // namespace faabric::util
