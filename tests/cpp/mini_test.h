// mini_test.h -- a few macros standing in for doctest (not installed here) so that the C++ drop-in tests read like
// the reference's test/cpp/*.cpp.
#pragma once
#include <cmath>
#include <cstdio>
#include <functional>
#include <string>
#include <vector>

namespace mini {
struct Case { std::string name; std::function<void()> fn; };
inline std::vector<Case>& cases() { static std::vector<Case> c; return c; }
inline int& failures() { static int f = 0; return f; }
inline int& checks() { static int c = 0; return c; }
struct Reg { Reg(const char* n, std::function<void()> f) { cases().push_back({n, std::move(f)}); } };
inline bool approx(double a, double b, double eps) { return std::fabs(a - b) <= eps * std::fmax(std::fabs(a), std::fabs(b)) || a == b; }
}  // namespace mini

#define MINI_CAT2(a, b) a##b
#define MINI_CAT(a, b) MINI_CAT2(a, b)
#define TEST_CASE(name)                                                    \
    static void MINI_CAT(mini_case_, __LINE__)();                            \
    static mini::Reg MINI_CAT(mini_reg_, __LINE__)(name, MINI_CAT(mini_case_, __LINE__)); \
    static void MINI_CAT(mini_case_, __LINE__)()
#define CHECK(cond)                                                                         \
    do {                                                                                    \
        ++mini::checks();                                                                   \
        if (!(cond)) { ++mini::failures(); std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)
#define CHECK_EQ(a, b) CHECK((a) == (b))
#define REQUIRE_EQ(a, b)                                                                    \
    do {                                                                                    \
        ++mini::checks();                                                                   \
        if (!((a) == (b))) { ++mini::failures(); std::printf("  FAILED (fatal) %s:%d: %s == %s\n", __FILE__, __LINE__, #a, #b); return; } \
    } while (0)
#define CHECK_APPROX(a, b, eps) CHECK(mini::approx((a), (b), (eps)))

inline int mini_main() {
    for (auto& c : mini::cases()) {
        const int before = mini::failures();
        c.fn();
        std::printf("[%s] %s\n", mini::failures() == before ? " ok " : "FAIL", c.name.c_str());
    }
    std::printf("%d checks, %d failures, %zu cases\n", mini::checks(), mini::failures(), mini::cases().size());
    return mini::failures() ? 1 : 0;
}
