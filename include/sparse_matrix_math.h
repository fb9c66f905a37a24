// sparse_matrix_math.h -- drop-in include name of vasil-pashov/sparse_matrix_math for the B200 build.
// Everything lives in smm_b200.hpp (C++17 surface, namespace SMM) over smm_b200.h (C ABI of libsmm_b200.so).
#pragma once
#include "smm_b200.hpp"
