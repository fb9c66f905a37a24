// oracle/tbb_shim/tbb/parallel_reduce.h -- TEST INFRASTRUCTURE ONLY (see blocked_range.h).
// parallel_deterministic_reduce (functional form): simple_partitioner semantics -- the range is split
// recursively while is_divisible(), each leaf is reduced by real_body(leaf, identity), and results are
// joined as reduction(left, right) up the split tree.  The value therefore depends only on the range and
// the grain size, never on the number of threads, which is what makes the reference's multithreaded dot
// product reproducible.  Note Value is deduced from the identity argument alone, as in oneTBB.
#pragma once
#include "blocked_range.h"
#ifdef _OPENMP
#include <omp.h>
#endif
namespace tbb {
namespace smm_shim_detail {
template <typename Range, typename Value, typename RealBody, typename Reduction>
Value reduce_tree(Range r, const Value& identity, const RealBody& body, const Reduction& red, int depth) {
    if (r.is_divisible()) {
        Range right(r, split());
        Value lv, rv;
        if (depth < 6 && r.size() > (1u << 15)) {
#pragma omp task shared(lv)
            lv = reduce_tree<Range, Value>(r, identity, body, red, depth + 1);
            rv = reduce_tree<Range, Value>(right, identity, body, red, depth + 1);
#pragma omp taskwait
        } else {
            lv = reduce_tree<Range, Value>(r, identity, body, red, depth + 1);
            rv = reduce_tree<Range, Value>(right, identity, body, red, depth + 1);
        }
        return red(lv, rv);
    }
    return body(r, identity);
}
}  // namespace smm_shim_detail
template <typename Range, typename Value, typename RealBody, typename Reduction>
Value parallel_deterministic_reduce(const Range& range, const Value& identity, const RealBody& real_body,
                                    const Reduction& reduction) {
    Value out = identity;
    if (range.size() <= (1u << 16)) {
        return smm_shim_detail::reduce_tree<Range, Value>(range, identity, real_body, reduction, 99);
    }
#pragma omp parallel
#pragma omp single
    out = smm_shim_detail::reduce_tree<Range, Value>(range, identity, real_body, reduction, 0);
    return out;
}
}  // namespace tbb
