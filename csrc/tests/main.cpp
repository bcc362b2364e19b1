#include "harness.h"

#include <faabric/util/config.h>
#include <faabric/util/crash.h>
#include <faabric/util/logging.h>
#include <faabric/util/testing.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#include <dirent.h>
#include <execinfo.h>
#include <signal.h>
#include <sys/syscall.h>
#include <unistd.h>

// Helper processes of tests that need a second process (test_threads_device.cpp)
int ipcMapChildMain(const char* hexHandle, const char* size);

namespace fbtest {
std::vector<TestCase>& registry()
{
    static std::vector<TestCase> r;
    return r;
}

long& assertionCount()
{
    static long n = 0;
    return n;
}
}

// ---- per-test watchdog: a wedged test (a device call that never returns)
// must not eat the whole session.  After FAABRIC_TEST_WATCHDOG_SECS (default
// 150) inside one test every thread prints its stack and the process exits.
namespace {
std::atomic<long> testStartedAtMs{ 0 };
std::atomic<const char*> currentTestName{ nullptr };

long nowMs()
{
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void dumpStackHandler(int)
{
    void* frames[48];
    int n = backtrace(frames, 48);
    char head[96];
    int len = snprintf(head, sizeof(head), "---- thread %ld ----\n", (long)syscall(SYS_gettid));
    (void)!write(2, head, (size_t)len);
    backtrace_symbols_fd(frames, n, 2);
}

void startWatchdog()
{
    long limitS = 150;
    if (const char* v = getenv("FAABRIC_TEST_WATCHDOG_SECS")) {
        limitS = atol(v);
    }
    if (limitS <= 0) {
        return;
    }
    struct sigaction sa{};
    sa.sa_handler = dumpStackHandler;
    sigaction(SIGUSR2, &sa, nullptr);
    std::thread([limitS] {
        for (;;) {
            std::this_thread::sleep_for(std::chrono::seconds(1));
            long t0 = testStartedAtMs.load();
            if (t0 != 0 && nowMs() - t0 > limitS * 1000) {
                const char* name = currentTestName.load();
                fprintf(stderr, "\n[watchdog] test '%s' exceeded %ld s: stacks of all threads follow\n", name ? name : "?", limitS);
                pid_t self = getpid();
                pid_t me = (pid_t)syscall(SYS_gettid);
                if (DIR* d = opendir("/proc/self/task")) {
                    while (dirent* e = readdir(d)) {
                        pid_t tid = (pid_t)atol(e->d_name);
                        if (tid > 0 && tid != me) {
                            syscall(SYS_tgkill, self, tid, SIGUSR2);
                            std::this_thread::sleep_for(std::chrono::milliseconds(30));
                        }
                    }
                    closedir(d);
                }
                fprintf(stderr, "[watchdog] giving up\n");
                _exit(3);
            }
        }
    }).detach();
}
}

int main(int argc, char** argv)
{
    faabric::util::setUpCrashHandler();
    faabric::util::setTestMode(true);
    faabric::util::initLogging();

    if (argc == 4 && !strcmp(argv[1], "--ipc-map-child")) {
        return ipcMapChildMain(argv[2], argv[3]);
    }
    bool list = false;
    std::string tag;
    std::vector<std::string> filters;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--list")) {
            list = true;
        } else if (!strcmp(argv[i], "--tag") && i + 1 < argc) {
            tag = argv[++i];
        } else {
            filters.emplace_back(argv[i]);
        }
    }

    startWatchdog();
    int passed = 0, failed = 0, skipped = 0;
    std::vector<std::string> failures;
    for (auto& t : fbtest::registry()) {
        bool selected = filters.empty();
        for (auto& f : filters) {
            selected = selected || t.name.find(f) != std::string::npos;
        }
        if (!tag.empty() && t.tags.find(tag) == std::string::npos) {
            selected = false;
        }
        if (!selected) {
            continue;
        }
        if (list) {
            printf("%s %s\n", t.name.c_str(), t.tags.c_str());
            continue;
        }
        auto t0 = std::chrono::steady_clock::now();
        std::string err;
        bool skip = false;
        currentTestName = t.name.c_str();
        testStartedAtMs = nowMs();
        try {
            t.fn();
        } catch (fbtest::TestSkipped& s) {
            skip = true;
            err = s.what();
        } catch (std::exception& e) {
            err = e.what();
            if (err.empty()) {
                err = "exception";
            }
        } catch (...) {
            err = "unknown exception";
        }
        testStartedAtMs = 0;
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (skip) {
            skipped++;
            printf("[ SKIP ] %s (%s)\n", t.name.c_str(), err.c_str());
        } else if (err.empty()) {
            passed++;
            printf("[  OK  ] %s (%.0f ms)\n", t.name.c_str(), ms);
        } else {
            failed++;
            failures.push_back(t.name);
            printf("[ FAIL ] %s (%.0f ms)\n         %s\n", t.name.c_str(), ms, err.c_str());
        }
        fflush(stdout);
    }
    if (!list) {
        printf("==== %d passed, %d failed, %d skipped, %ld assertions ====\n",
               passed,
               failed,
               skipped,
               fbtest::assertionCount());
        for (auto& f : failures) {
            printf("  failed: %s\n", f.c_str());
        }
    }
    return failed == 0 ? 0 : 1;
}
