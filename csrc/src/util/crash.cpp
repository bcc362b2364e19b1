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

static void crashHandler(int sig, siginfo_t* info, void*) noexcept
{
    if (info != nullptr && (sig == SIGSEGV || sig == SIGBUS)) {
        fprintf(stderr, "Caught fatal signal %d (fault address %p)\n", sig, info->si_addr);
    } else {
        fprintf(stderr, "Caught fatal signal %d\n", sig);
    }
    printStackTrace();
    if (sig != TEST_SIGNAL) {
        ::signal(sig, SIG_DFL);
        ::raise(sig);
        ::_exit(1);
    }
}

void handleCrash(int sig)
{
    crashHandler(sig, nullptr, nullptr);
}

static void installHandler(int s)
{
    struct sigaction sa{};
    sa.sa_sigaction = crashHandler;
    sigemptyset(&sa.sa_mask);
    // Run on the alternate stack so stack overflows still get a trace
    sa.sa_flags = SA_ONSTACK | SA_SIGINFO;
    if (::sigaction(s, &sa, nullptr) != 0) {
        SPDLOG_WARN("Could not install crash handler for signal {}", s);
    }
}

void setUpCrashHandler(int sig)
{
    // The segfault dirty tracker installs itself over SIGSEGV later and
    // falls back to this handler for faults outside tracked memory
    int signals[] = { SIGSEGV, SIGABRT, SIGILL, SIGFPE, SIGBUS };
    if (sig >= 0) {
        if (sig == TEST_SIGNAL) {
            crashHandler(sig, nullptr, nullptr);
            return;
        }
        installHandler(sig);
        return;
    }
    static thread_local char altStack[64 * 1024];
    stack_t ss{};
    ss.ss_sp = altStack;
    ss.ss_size = sizeof(altStack);
    ss.ss_flags = 0;
    ::sigaltstack(&ss, nullptr);
    for (int s : signals) {
        installHandler(s);
    }
}

} // namespace faabric::util
