// oracle/tbb_shim/tbb/blocked_range.h -- TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for oneTBB (not installed in this image, no network) so that the UNMODIFIED reference
// header can be compiled with -DSMM_MULTITHREADING into oracle/_ref/.  Written from oneTBB's documented
// semantics: blocked_range is divisible while size() > grainsize and splits at begin + size()/2.
#pragma once
#include <cstddef>
namespace tbb {
struct split {};
template <typename Value>
class blocked_range {
public:
    using const_iterator = Value;
    using size_type = std::size_t;
    blocked_range(Value b, Value e, size_type grain = 1) : my_end(e), my_begin(b), my_grainsize(grain) {}
    blocked_range(blocked_range& r, split) : my_end(r.my_end), my_begin(do_split(r)), my_grainsize(r.my_grainsize) {}
    const_iterator begin() const { return my_begin; }
    const_iterator end() const { return my_end; }
    size_type size() const { return size_type(my_end - my_begin); }
    size_type grainsize() const { return my_grainsize; }
    bool empty() const { return !(my_begin < my_end); }
    bool is_divisible() const { return my_grainsize < size(); }
private:
    Value my_end, my_begin;
    size_type my_grainsize;
    static Value do_split(blocked_range& r) {
        Value middle = r.my_begin + (r.my_end - r.my_begin) / 2u;
        r.my_end = middle;
        return middle;
    }
};
}  // namespace tbb
