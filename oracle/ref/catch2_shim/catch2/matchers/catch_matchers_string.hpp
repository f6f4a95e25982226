// oracle/ref/catch2_shim -- TEST INFRASTRUCTURE ONLY (see catch2/catch_test_macros.hpp)
#pragma once
#include <catch2/matchers/catch_matchers.hpp>
