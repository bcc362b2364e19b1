#include <faabric/util/crash.h>
#include <faabric/util/logging.h>

#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#include <unistd.h>

namespace faabric::util {

constexpr int TEST_SIGNAL = 12341234;
constexpr int MAX_FRAMES = 100;

void printStackTrace(void* contextR)
{
    void* frames[MAX_FRAMES];
    int n = ::backtrace(frames, MAX_FRAMES);
    fprintf(stderr, "Stack trace (%d frames):\n", n);
    ::backtrace_symbols_fd(frames, n, STDERR_FILENO);
}

static void crashHandler(int sig) noexcept
{
    fprintf(stderr, "Caught fatal signal %d\n", sig);
    printStackTrace();
    if (sig != TEST_SIGNAL) {
        ::signal(sig, SIG_DFL);
        ::raise(sig);
        ::_exit(1);
    }
}

void setUpCrashHandler(int sig)
{
    // SIGSEGV is deliberately absent: the segfault dirty tracker owns it
    int signals[] = { SIGABRT, SIGILL, SIGFPE };
    if (sig >= 0) {
        if (sig == TEST_SIGNAL) {
            crashHandler(sig);
            return;
        }
        ::signal(sig, crashHandler);
        return;
    }
    for (int s : signals) {
        if (::signal(s, crashHandler) == SIG_ERR) {
            SPDLOG_WARN("Could not install crash handler for signal {}", s);
        }
    }
}

} // namespace faabric::util
