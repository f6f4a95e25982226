// oracle/ref/catch2_shim — TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the handful of Catch2 v3 macros and matchers the reference's unit tests use (test/*.cc: TEST_CASE,
// REQUIRE, CHECK, REQUIRE_FALSE, REQUIRE_THAT / CHECK_THAT, REQUIRE_THROWS_WITH, REQUIRE_THROWS_MATCHES, SKIP, BENCHMARK;
// matchers WithinAbs, Equals, Approx, Message, ContainsSubstring).  Catch2 itself is not installed here and cannot be
// fetched; with this header the reference's OWN, unmodified test sources compile where they lie and run against the
// façade (oracle/ref/Makefile), on the CPU backend of the C ABI and on the B200.
//
// Runner: every TEST_CASE whose tags do not contain "[integration]" (those need downloaded model weights) or "[!benchmark]";
// exit code = number of failed test cases; one line per case on stdout.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

namespace Catch {

struct TestFailure : std::exception {
    std::string message;
    explicit TestFailure(std::string m)
    : message(std::move(m))
    {}
    const char*
    what() const noexcept override
    {
        return message.c_str();
    }
};
struct TestSkipped : std::exception {};

struct TestCase {
    const char* name;
    const char* tags;
    void (*fn)();
};
inline std::vector<TestCase>&
registry()
{
    static std::vector<TestCase> r;
    return r;
}
struct Registrar {
    Registrar(const char* name, const char* tags, void (*fn)()) { registry().push_back({name, tags, fn}); }
};
inline int&
soft_failures()
{
    static int n = 0;
    return n;
}
inline std::string
location(const char* file, int line)
{
    return std::string(file) + ":" + std::to_string(line);
}
template <typename T>
std::string
stringify(const T& v)
{
    if constexpr (requires(std::ostringstream& os, const T& x) { os << x; }) {
        std::ostringstream os;
        os << v;
        return os.str();
    } else {
        return "<value>";
    }
}
inline void
require(bool ok, const char* expr, const char* file, int line, bool hard)
{
    if (ok) {
        return;
    }
    const std::string msg = location(file, line) + ": FAILED: " + expr;
    if (hard) {
        throw TestFailure(msg);
    }
    std::printf("    %s\n", msg.c_str());
    soft_failures()++;
}
template <typename V, typename M>
void
require_that(const V& value, const M& matcher, const char* expr, const char* file, int line, bool hard)
{
    if (matcher.match(value)) {
        return;
    }
    const std::string msg = location(file, line) + ": FAILED: " + expr + " (" + matcher.describe() + ")";
    if (hard) {
        throw TestFailure(msg);
    }
    std::printf("    %s\n", msg.c_str());
    soft_failures()++;
}

namespace Benchmark {
struct Benchmark {
    std::string name;
    explicit operator bool() const { return std::getenv("MC_REFTEST_BENCHMARKS") != nullptr; }
    template <typename F>
    Benchmark&
    operator=(F&& f)
    {
        if constexpr (std::is_invocable_v<F, int>) {
            f(0);
        } else {
            f();
        }
        return *this;
    }
};
} // namespace Benchmark

} // namespace Catch

#define CATCH_SHIM_CAT2(a, b) a##b
#define CATCH_SHIM_CAT(a, b) CATCH_SHIM_CAT2(a, b)
#define CATCH_SHIM_TEST(fn, ...)                                   \
    static void fn();                                              \
    static ::Catch::Registrar CATCH_SHIM_CAT(fn, _reg)(CATCH_SHIM_FIRST(__VA_ARGS__, ""), CATCH_SHIM_SECOND(__VA_ARGS__, "", ""), &fn); \
    static void fn()
#define CATCH_SHIM_FIRST(a, ...) a
#define CATCH_SHIM_SECOND(a, b, ...) b
#define TEST_CASE(...) CATCH_SHIM_TEST(CATCH_SHIM_CAT(catch_shim_test_, __LINE__), __VA_ARGS__)

#define REQUIRE(...) ::Catch::require(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__, true)
#define CHECK(...) ::Catch::require(static_cast<bool>(__VA_ARGS__), #__VA_ARGS__, __FILE__, __LINE__, false)
#define REQUIRE_FALSE(...) ::Catch::require(!static_cast<bool>(__VA_ARGS__), "!(" #__VA_ARGS__ ")", __FILE__, __LINE__, true)
#define CHECK_FALSE(...) ::Catch::require(!static_cast<bool>(__VA_ARGS__), "!(" #__VA_ARGS__ ")", __FILE__, __LINE__, false)
#define REQUIRE_THAT(value, matcher) ::Catch::require_that((value), (matcher), #value ", " #matcher, __FILE__, __LINE__, true)
#define CHECK_THAT(value, matcher) ::Catch::require_that((value), (matcher), #value ", " #matcher, __FILE__, __LINE__, false)
#define SKIP(...) throw ::Catch::TestSkipped()
#define SUCCEED(...) ((void)0)
#define INFO(...) ((void)0)

#define REQUIRE_THROWS(...)                                                                                     \
    do {                                                                                                        \
        bool catch_shim_thrown = false;                                                                         \
        try {                                                                                                   \
            static_cast<void>(__VA_ARGS__);                                                                     \
        } catch (...) {                                                                                         \
            catch_shim_thrown = true;                                                                           \
        }                                                                                                       \
        ::Catch::require(catch_shim_thrown, "throws: " #__VA_ARGS__, __FILE__, __LINE__, true);                \
    } while (0)

// REQUIRE_THROWS_WITH(expr, "exact message" | string matcher)
#define REQUIRE_THROWS_WITH(expr, matcher)                                                                      \
    do {                                                                                                        \
        bool catch_shim_thrown = false;                                                                         \
        try {                                                                                                   \
            static_cast<void>(expr);                                                                            \
        } catch (const std::exception& e) {                                                                     \
            catch_shim_thrown = true;                                                                           \
            ::Catch::require(::Catch::message_matches(std::string(e.what()), matcher), "message of " #expr " matches " #matcher, __FILE__, __LINE__, true); \
        } catch (...) {                                                                                         \
            catch_shim_thrown = true;                                                                           \
        }                                                                                                       \
        ::Catch::require(catch_shim_thrown, "throws: " #expr, __FILE__, __LINE__, true);                        \
    } while (0)

#define REQUIRE_THROWS_MATCHES(expr, exception_type, matcher)                                                   \
    do {                                                                                                        \
        bool catch_shim_thrown = false;                                                                         \
        try {                                                                                                   \
            static_cast<void>(expr);                                                                            \
        } catch (const exception_type& e) {                                                                     \
            catch_shim_thrown = true;                                                                           \
            ::Catch::require((matcher).match(e), "exception of " #expr " matches " #matcher, __FILE__, __LINE__, true); \
        } catch (...) {                                                                                         \
        }                                                                                                       \
        ::Catch::require(catch_shim_thrown, "throws " #exception_type ": " #expr, __FILE__, __LINE__, true);    \
    } while (0)

namespace Catch {
inline bool
message_matches(const std::string& what, const std::string& expect)
{
    return what == expect;
}
inline bool
message_matches(const std::string& what, const char* expect)
{
    return what == expect;
}
template <typename M>
bool
message_matches(const std::string& what, const M& matcher)
{
    return matcher.match(what);
}
} // namespace Catch

#ifndef CATCH_SHIM_NO_MAIN
#include <csetjmp>
#include <csignal>
// Integer division by zero returns 0 on the reference's only target (arm64) and raises SIGFPE on x86-64; a test case that does it
// (test/test_tensor.cc "Tensor empty": the iterator of a zero-sized tensor) is reported as `trapped`, separately from failures.
static sigjmp_buf catch_shim_trap;
static void
catch_shim_on_sigfpe(int)
{
    siglongjmp(catch_shim_trap, 1);
}

int
main(int argc, char** argv)
{
    struct sigaction catch_shim_sa;
    std::memset(&catch_shim_sa, 0, sizeof(catch_shim_sa));
    catch_shim_sa.sa_handler = catch_shim_on_sigfpe;
    catch_shim_sa.sa_flags = SA_NODEFER;
    sigaction(SIGFPE, &catch_shim_sa, nullptr);
    int trapped = 0;
    bool with_integration = std::getenv("MC_REFTEST_INTEGRATION") != nullptr;
    const char* only = nullptr;
    for (int i = 1; i < argc; i++) {
        if (std::strncmp(argv[i], "--", 2) != 0 && argv[i][0] != '~' && argv[i][0] != '[') {
            only = argv[i];
        }
    }
    int failed = 0, passed = 0, skipped = 0;
    for (const auto& tc : Catch::registry()) {
        const std::string tags = tc.tags ? tc.tags : "";
        if ((!with_integration && tags.find("[integration]") != std::string::npos) || tags.find("[!benchmark]") != std::string::npos ||
            (only != nullptr && std::string(tc.name).find(only) == std::string::npos)) {
            skipped++;
            continue;
        }
        const int soft_before = Catch::soft_failures();
        if (sigsetjmp(catch_shim_trap, 1) != 0) {
            std::printf("trapped %s %s (SIGFPE: integer division by zero, which arm64 answers with 0)\n", tc.name, tags.c_str());
            trapped++;
            continue;
        }
        try {
            tc.fn();
            if (Catch::soft_failures() != soft_before) {
                std::printf("FAILED  %s %s\n", tc.name, tags.c_str());
                failed++;
            } else {
                std::printf("passed  %s %s\n", tc.name, tags.c_str());
                passed++;
            }
        } catch (const Catch::TestSkipped&) {
            std::printf("skipped %s %s\n", tc.name, tags.c_str());
            skipped++;
        } catch (const std::exception& e) {
            std::printf("FAILED  %s %s\n    %s\n", tc.name, tags.c_str(), e.what());
            failed++;
        } catch (...) {
            std::printf("FAILED  %s %s\n    unknown exception\n", tc.name, tags.c_str());
            failed++;
        }
        std::fflush(stdout);
    }
    std::printf("== %d passed, %d failed, %d skipped, %d trapped\n", passed, failed, skipped, trapped);
    return failed;
}
#endif
