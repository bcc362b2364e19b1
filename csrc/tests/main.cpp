#include "harness.h"

#include <faabric/util/config.h>
#include <faabric/util/crash.h>
#include <faabric/util/logging.h>
#include <faabric/util/testing.h>

#include <chrono>
#include <cstdio>
#include <cstring>

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
