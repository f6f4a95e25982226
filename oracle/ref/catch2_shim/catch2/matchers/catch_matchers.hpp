// oracle/ref/catch2_shim — TEST INFRASTRUCTURE ONLY: the matchers used by the reference's tests (see catch_test_macros.hpp).
#pragma once

#include <cmath>
#include <string>
#include <vector>

#include <catch2/catch_test_macros.hpp>

namespace Catch {
namespace Matchers {

struct WithinAbsMatcher {
    double target, margin;
    template <typename T>
    bool
    match(const T& v) const
    {
        return std::fabs(double(v) - target) <= margin;
    }
    std::string
    describe() const
    {
        return "is within " + std::to_string(margin) + " of " + std::to_string(target);
    }
};
template <typename T, typename M>
WithinAbsMatcher
WithinAbs(const T& target, const M& margin)
{
    return WithinAbsMatcher{double(target), double(margin)};
}

template <typename T> struct VectorEqualsMatcher {
    const std::vector<T>& expect;
    bool
    match(const std::vector<T>& v) const
    {
        return v == expect;
    }
    std::string
    describe() const
    {
        return "equals the expected vector";
    }
};
template <typename T>
VectorEqualsMatcher<T>
Equals(const std::vector<T>& expect)
{
    return VectorEqualsMatcher<T>{expect};
}

struct StringEqualsMatcher {
    std::string expect;
    bool
    match(const std::string& v) const
    {
        return v == expect;
    }
    std::string
    describe() const
    {
        return "equals \"" + expect + "\"";
    }
};
inline StringEqualsMatcher
Equals(const std::string& expect)
{
    return StringEqualsMatcher{expect};
}
inline StringEqualsMatcher
Equals(const char* expect)
{
    return StringEqualsMatcher{expect};
}

// Catch::Approx semantics: |a - b| <= margin, or within epsilon (100 * float epsilon) relative to the larger magnitude
template <typename T> struct VectorApproxMatcher {
    const std::vector<T>& expect;
    double margin_ = 0.0, epsilon_ = 100.0 * 1.1920929e-07;
    VectorApproxMatcher&
    margin(double m)
    {
        margin_ = m;
        return *this;
    }
    VectorApproxMatcher&
    epsilon(double e)
    {
        epsilon_ = e;
        return *this;
    }
    bool
    match(const std::vector<T>& v) const
    {
        if (v.size() != expect.size()) {
            return false;
        }
        for (std::size_t i = 0; i < v.size(); i++) {
            const double a = double(v[i]), b = double(expect[i]);
            const double d = std::fabs(a - b);
            if (!(d <= margin_ || d <= epsilon_ * (std::isinf(b) ? 0.0 : std::fabs(b)))) {
                return false;
            }
        }
        return true;
    }
    std::string
    describe() const
    {
        return "is approximately the expected vector";
    }
};
template <typename T>
VectorApproxMatcher<T>
Approx(const std::vector<T>& expect)
{
    return VectorApproxMatcher<T>{expect};
}

struct ExceptionMessageMatcher {
    std::string expect;
    bool
    match(const std::exception& e) const
    {
        return expect == e.what();
    }
    std::string
    describe() const
    {
        return "exception message is \"" + expect + "\"";
    }
};
inline ExceptionMessageMatcher
Message(const std::string& expect)
{
    return ExceptionMessageMatcher{expect};
}

struct ContainsSubstringMatcher {
    std::string needle;
    bool
    match(const std::string& v) const
    {
        return v.find(needle) != std::string::npos;
    }
    std::string
    describe() const
    {
        return "contains \"" + needle + "\"";
    }
};
inline ContainsSubstringMatcher
ContainsSubstring(const std::string& needle)
{
    return ContainsSubstringMatcher{needle};
}

} // namespace Matchers
} // namespace Catch
