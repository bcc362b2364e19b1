// Minimal logger with the spdlog-style macro surface the reference uses
// (include/faabric/util/logging.h).  Pattern: [HH:MM:SS.mmm] [tid] [L] msg.
// Format strings use {} placeholders.
#pragma once

#include <sstream>
#include <string>
#include <string_view>

namespace faabric::util {

enum class LogLevel : int
{
    trace = 0,
    debug = 1,
    info = 2,
    warn = 3,
    err = 4,
    critical = 5,
    off = 6
};

void initLogging();

LogLevel getLogLevel();

void setLogLevel(LogLevel level);

void setLogLevel(const std::string& name);

void logLine(LogLevel level, const std::string& msg);

namespace detail {
inline void fmtInto(std::ostringstream& os, std::string_view f)
{
    os << f;
}

template<typename T, typename... Rest>
void fmtInto(std::ostringstream& os,
             std::string_view f,
             const T& v,
             const Rest&... rest)
{
    size_t pos = f.find("{}");
    if (pos == std::string_view::npos) {
        os << f;
        return;
    }
    os << f.substr(0, pos) << v;
    fmtInto(os, f.substr(pos + 2), rest...);
}
}

template<typename... Args>
std::string format(std::string_view f, const Args&... args)
{
    std::ostringstream os;
    detail::fmtInto(os, f, args...);
    return os.str();
}

template<typename... Args>
void logFmt(LogLevel level, std::string_view f, const Args&... args)
{
    if ((int)level < (int)getLogLevel()) {
        return;
    }
    logLine(level, format(f, args...));
}

} // namespace faabric::util

// Compile-time floor: trace/debug compiled out unless FAABRIC_LOG_DEBUG is set
#ifdef FAABRIC_LOG_DEBUG
#define SPDLOG_TRACE(...)                                                      \
    faabric::util::logFmt(faabric::util::LogLevel::trace, __VA_ARGS__)
#define SPDLOG_DEBUG(...)                                                      \
    faabric::util::logFmt(faabric::util::LogLevel::debug, __VA_ARGS__)
#else
#define SPDLOG_TRACE(...) (void)0
#define SPDLOG_DEBUG(...)                                                      \
    faabric::util::logFmt(faabric::util::LogLevel::debug, __VA_ARGS__)
#endif
#define SPDLOG_INFO(...)                                                       \
    faabric::util::logFmt(faabric::util::LogLevel::info, __VA_ARGS__)
#define SPDLOG_WARN(...)                                                       \
    faabric::util::logFmt(faabric::util::LogLevel::warn, __VA_ARGS__)
#define SPDLOG_ERROR(...)                                                      \
    faabric::util::logFmt(faabric::util::LogLevel::err, __VA_ARGS__)
#define SPDLOG_CRITICAL(...)                                                   \
    faabric::util::logFmt(faabric::util::LogLevel::critical, __VA_ARGS__)
