// Minimal in-tree C++ test harness (no Catch2 in the image).  TEST_CASE
// registers a function; REQUIRE* throw a TestFailure which the runner reports.
// Usage: faabric_tests [--list] [--tag TAG] [substring-filter ...]
#pragma once

#include <cmath>
#include <functional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace fbtest {

struct TestFailure : public std::runtime_error
{
    using std::runtime_error::runtime_error;
};

struct TestSkipped : public std::runtime_error
{
    using std::runtime_error::runtime_error;
};

struct TestCase
{
    std::string name;
    std::string tags;
    std::function<void()> fn;
};

std::vector<TestCase>& registry();

struct Registrar
{
    Registrar(const char* name, const char* tags, std::function<void()> fn)
    {
        registry().push_back({ name, tags, std::move(fn) });
    }
};

template<typename T>
std::string show(const T& v)
{
    if constexpr (requires(std::ostringstream& o, const T& x) { o << x; }) {
        std::ostringstream o;
        o << v;
        return o.str();
    } else {
        return "<?>";
    }
}

inline std::string show(const std::string& s)
{
    return "\"" + s + "\"";
}

[[noreturn]] inline void fail(const char* file, int line, const std::string& what)
{
    throw TestFailure(std::string(file) + ":" + std::to_string(line) + ": " + what);
}

// Number of assertions evaluated (reported in the summary)
long& assertionCount();

}

#define FB_CAT2(a, b) a##b
#define FB_CAT(a, b) FB_CAT2(a, b)

#define TEST_CASE(name, tags)                                                  \
    static void FB_CAT(fbtest_fn_, __LINE__)();                                \
    static fbtest::Registrar FB_CAT(fbtest_reg_, __LINE__)(name, tags, FB_CAT(fbtest_fn_, __LINE__)); \
    static void FB_CAT(fbtest_fn_, __LINE__)()

#define REQUIRE(cond)                                                          \
    do {                                                                       \
        fbtest::assertionCount()++;                                            \
        if (!(cond)) {                                                         \
            fbtest::fail(__FILE__, __LINE__, "REQUIRE(" #cond ") failed");     \
        }                                                                      \
    } while (0)

#define REQUIRE_EQ(a, b)                                                       \
    do {                                                                       \
        fbtest::assertionCount()++;                                            \
        auto _a = (a);                                                         \
        auto _b = (b);                                                         \
        if (!(_a == _b)) {                                                     \
            fbtest::fail(__FILE__, __LINE__, "REQUIRE_EQ(" #a ", " #b ") failed: " + fbtest::show(_a) + " != " + fbtest::show(_b)); \
        }                                                                      \
    } while (0)

#define REQUIRE_NEAR(a, b, eps)                                                \
    do {                                                                       \
        fbtest::assertionCount()++;                                            \
        double _a = (double)(a);                                               \
        double _b = (double)(b);                                               \
        if (std::fabs(_a - _b) > (eps)) {                                      \
            fbtest::fail(__FILE__, __LINE__, "REQUIRE_NEAR(" #a ", " #b ") failed: " + std::to_string(_a) + " vs " + std::to_string(_b)); \
        }                                                                      \
    } while (0)

#define REQUIRE_THROWS(expr)                                                   \
    do {                                                                       \
        fbtest::assertionCount()++;                                            \
        bool _threw = false;                                                   \
        try {                                                                  \
            (void)(expr);                                                      \
        } catch (fbtest::TestFailure&) {                                       \
            throw;                                                             \
        } catch (...) {                                                        \
            _threw = true;                                                     \
        }                                                                      \
        if (!_threw) {                                                         \
            fbtest::fail(__FILE__, __LINE__, "REQUIRE_THROWS(" #expr ") did not throw"); \
        }                                                                      \
    } while (0)

#define REQUIRE_NOTHROW(expr)                                                  \
    do {                                                                       \
        fbtest::assertionCount()++;                                            \
        try {                                                                  \
            (void)(expr);                                                      \
        } catch (fbtest::TestFailure&) {                                       \
            throw;                                                             \
        } catch (std::exception & _e) {                                        \
            fbtest::fail(__FILE__, __LINE__, std::string("REQUIRE_NOTHROW(" #expr ") threw: ") + _e.what()); \
        }                                                                      \
    } while (0)

#define SKIP_TEST(why) throw fbtest::TestSkipped(why)
