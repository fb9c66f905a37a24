// oracle/tbb_shim/tbb/parallel_for.h -- TEST INFRASTRUCTURE ONLY (see blocked_range.h).
// parallel_for over a blocked_range: the body is applied to disjoint sub-ranges that cover the range.
// Every loop the reference passes here is element-wise / row-wise, so the chunking cannot change results.
#pragma once
#include "blocked_range.h"
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif
namespace tbb {
inline long long& smm_shim_parallel_for_calls() { static long long n = 0; return n; }
template <typename Value, typename Body>
void parallel_for(const blocked_range<Value>& range, const Body& body) {
    ++smm_shim_parallel_for_calls();
    const long long b = range.begin(), e = range.end();
    if (e <= b) return;
    const long long chunk = 2048;
    const long long nchunks = (e - b + chunk - 1) / chunk;
#pragma omp parallel for schedule(static)
    for (long long c = 0; c < nchunks; ++c) {
        const long long lo = b + c * chunk, hi = std::min(e, lo + chunk);
        body(blocked_range<Value>(Value(lo), Value(hi)));
    }
}
}  // namespace tbb
