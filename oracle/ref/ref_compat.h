// oracle/ref/ref_compat.h — TEST INFRASTRUCTURE ONLY.  Pre-included (-include) when the reference's sources are compiled with
// g++ / libstdc++ instead of Apple clang / libc++; nothing here changes what the reference computes.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <optional>
#include <string>

namespace std {
// libc++ exposes the C float overloads in namespace std (test/test_kernel_embedding.cc:86 calls std::powf); libstdc++ 13 does not
using ::cosf;
using ::powf;
using ::sinf;
} // namespace std
