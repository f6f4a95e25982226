// oracle/ref/catch2_shim -- TEST INFRASTRUCTURE ONLY.  BENCHMARK("name") { body }; runs its body once when
// MC_REFTEST_BENCHMARKS is set and is skipped otherwise (the reference's ctest line passes --skip-benchmarks).
#pragma once
#include <catch2/catch_test_macros.hpp>

#define BENCHMARK(...) if (::Catch::Benchmark::Benchmark catch_shim_benchmark{std::string(__VA_ARGS__)}) catch_shim_benchmark = [&]
