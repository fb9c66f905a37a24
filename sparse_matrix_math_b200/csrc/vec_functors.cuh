// vec_functors.cuh -- the element-wise updates of the Krylov solvers as functors (one per loop of the reference), the
// parameter block of the kernels that run them, and the halo store of the multi-GPU CG.  Shared by vecops.cu (one kernel per
// update) and the persistent CG iteration in spmv.cu.  Internal linkage: include inside the translation unit's own code.
#pragma once
#include <type_traits>

#include "epilogue.cuh"
#include "smm_internal.cuh"

namespace {

constexpr int VEC_THREADS = 256;
constexpr int VEC_CTAS_PER_SM = 8;

struct VecParams {
    long long n;
    const float* in[5];
    float* out[3];
    SolveState* state;
    int finish;
    float* partials;
    size_t partials_stride;
    unsigned int* ticket;
    const HaloPushDev* halo;   // multi-GPU: boundary entries of out[0] are also stored into the peers' extended vectors
};

// out[0][e .. e + cnt) has just been computed: the part that lies in a send segment goes to the peer as well (P2P stores)
__device__ __forceinline__ void halo_store(const HaloSeg* segs, const int nsegs, const long long e, const float* v, const int cnt, bool& pushed) {
    for (int s = 0; s < nsegs; ++s) {
        const long long b = segs[s].begin, len = segs[s].len;
        if (e + cnt <= b || e >= b + len) continue;
        float* dst = segs[s].dst + (e - b);
        if (cnt == 4 && e >= b && e + 4 <= b + len && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int c = 0; c < cnt; ++c) if (e + c >= b && e + c < b + len) dst[c] = v[c];
        }
        pushed = true;
    }
}

// F::OWED (optional): the kernel has work to do after the stopping test has fired (see FCgPX)
template <class F, class = void> struct vec_owed : std::false_type {};
template <class F> struct vec_owed<F, std::void_t<decltype(F::OWED)>> : std::integral_constant<bool, F::OWED> {};

// ---- functors: NIN inputs, NOUT outputs, NRED reductions; sc = scalars read once per thread -------------------
struct Scal { float a, b, c; };

// CG  x = fma(alpha,p,x); r = fma(-alpha,Ap,r); t0 = r.r            (H:2363-2375)   in: x p r Ap  out: x r
struct FCgXR {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 4, NOUT = 2, NRED = 1;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float* red) {
        out[0] = smm_fma2(sc.a, in[1], in[0]);
        const float r = smm_fma2(-sc.a, in[3], in[2]);
        out[1] = r;
        red[0] = fmaf(r, r, red[0]);
    }
};
// CG  p = fma(beta,p,r)                                              (H:2385-2393)   in: p r  out: p
struct FCgP {
    static constexpr bool HALO_OK = true;
    static constexpr int NIN = 2, NOUT = 1, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) { out[0] = smm_fma2(sc.a, in[0], in[1]); }
};
// The same iteration in two passes that read p once instead of twice (32 n bytes instead of 36 n): the x update does not feed
// the stopping test, so it waits for the pass that reads p anyway.
// CG  r = fma(-alpha,Ap,r); t0 = r.r                                 (H:2366-2375)   in: r Ap  out: r
struct FCgR {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 2, NOUT = 1, NRED = 1;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float* red) {
        const float r = smm_fma2(-sc.a, in[1], in[0]);
        out[0] = r;
        red[0] = fmaf(r, r, red[0]);
    }
};
// CG  x = fma(alpha,p,x) with the p of this iteration (H:2363-2365), then p = fma(beta,p,r) (H:2385-2393)
//                                                                     in: p r x  out: p x
// OWED: when the r update's stopping test has fired (state->done with state->x_owed) the kernel still runs once, for the x
// update alone -- the reference updates x before it tests (H:2363-2379) -- and leaves p as it is (sc.c != 0).
struct FCgPX {
    static constexpr bool HALO_OK = true;
    static constexpr bool OWED = true;
    static constexpr int NIN = 3, NOUT = 2, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, s->alpha, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) {
        out[1] = smm_fma2(sc.b, in[0], in[2]);
        out[0] = sc.c != 0.f ? in[0] : smm_fma2(sc.a, in[0], in[1]);
    }
};
// BiCGSymmetric  x += alpha*p; r -= alpha*ap; t0 = r.r               (H:2061-2075)   in: x p r ap  out: x r
struct FBsXR {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 4, NOUT = 2, NRED = 1;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float* red) {
        out[0] = __fadd_rn(in[0], __fmul_rn(sc.a, in[1]));
        const float r = __fsub_rn(in[2], __fmul_rn(sc.a, in[3]));
        out[1] = r;
        red[0] = fmaf(r, r, red[0]);
    }
};
// BiCGSymmetric in two passes that read p once, like FCgR / FCgPX:  r -= alpha*ap; t0 = r.r   (H:2068-2075)   in: r ap  out: r
struct FBsR {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 2, NOUT = 1, NRED = 1;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float* red) {
        const float r = __fsub_rn(in[0], __fmul_rn(sc.a, in[1]));
        out[0] = r;
        red[0] = fmaf(r, r, red[0]);
    }
};
// BiCGSymmetric  x += alpha*p (H:2061-2067), then p = r + beta*p (H:2084-2092)        in: p r x  out: p x   (OWED: see FCgPX)
struct FBsPX {
    static constexpr bool HALO_OK = false;
    static constexpr bool OWED = true;
    static constexpr int NIN = 3, NOUT = 2, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, s->alpha, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) {
        out[1] = __fadd_rn(in[2], __fmul_rn(sc.b, in[0]));
        out[0] = sc.c != 0.f ? in[0] : __fadd_rn(in[1], __fmul_rn(sc.a, in[0]));
    }
};
// BiCGSymmetric  p = r + beta*p                                      (H:2084-2092)   in: p r  out: p
struct FBsP {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 2, NOUT = 1, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) { out[0] = __fadd_rn(in[1], __fmul_rn(sc.a, in[0])); }
};
// CGS  q = fma(-alpha,ap,u); auq = alpha*(u+q); x = x + auq          (H:2137-2149)   in: ap u x  out: q auq x
struct FCgsQX {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 3, NOUT = 3, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) {
        const float q = smm_fma2(-sc.a, in[0], in[1]);
        const float auq = __fmul_rn(sc.a, __fadd_rn(in[1], q));
        out[0] = q;
        out[1] = auq;
        out[2] = __fadd_rn(in[2], auq);
    }
};
// CGS  u = fma(beta,q,r); p = fma(beta, fma(beta,p,q), u)            (H:2157-2167)   in: q r p  out: u p
struct FCgsUP {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 3, NOUT = 2, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) {
        const float u = smm_fma2(sc.a, in[0], in[1]);
        out[0] = u;
        out[1] = smm_fma2(sc.a, smm_fma2(sc.a, in[2], in[0]), u);
    }
};
// BiCGStab  s = fma(-alpha,ap,r)                                     (H:2245-2247)   in: ap r  out: s
struct FStabS {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 2, NOUT = 1, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, 0.f, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) { out[0] = smm_fma2(-sc.a, in[0], in[1]); }
};
// BiCGStab  x = fma(alpha,p,fma(omega,s,x)); r = fma(-omega,as,s); t0 = r.r; t1 = r.r0   (H:2263-2269)
//                                                                     in: x p s as r0  out: x r
struct FStabXR {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 5, NOUT = 2, NRED = 2;
    static __device__ Scal scal(const SolveState* s) { return {s->alpha, s->omega, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float* red) {
        out[0] = smm_fma2(sc.a, in[1], smm_fma2(sc.b, in[2], in[0]));
        const float r = smm_fma2(-sc.b, in[3], in[2]);
        out[1] = r;
        red[0] = fmaf(r, r, red[0]);
        red[1] = fmaf(r, in[4], red[1]);
    }
};
// BiCGStab  p = fma(beta, fma(-omega,ap,p), r)                       (H:2272-2274)   in: p ap r  out: p
struct FStabP {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 3, NOUT = 1, NRED = 0;
    static __device__ Scal scal(const SolveState* s) { return {s->beta, s->omega, 0.f}; }
    static __device__ void apply(const Scal& sc, const float* in, float* out, float*) {
        out[0] = smm_fma2(sc.a, smm_fma2(-sc.b, in[1], in[0]), in[2]);
    }
};
// dot products: t0 = a.b, t1 = a.a                                                    in: a b
struct FDot2 {
    static constexpr bool HALO_OK = false;
    static constexpr int NIN = 2, NOUT = 0, NRED = 2;
    static __device__ Scal scal(const SolveState*) { return {0.f, 0.f, 0.f}; }
    static __device__ void apply(const Scal&, const float* in, float*, float* red) {
        red[0] = fmaf(in[0], in[1], red[0]);
        red[1] = fmaf(in[0], in[0], red[1]);
    }
};
// copies: out0 = out1 = out2 = in0 (r0 = p = r after the preconditioned start, H:2221-2227)   in: a
struct FCopy3 {
    static constexpr bool HALO_OK = true;
    static constexpr int NIN = 1, NOUT = 3, NRED = 0;
    static __device__ Scal scal(const SolveState*) { return {0.f, 0.f, 0.f}; }
    static __device__ void apply(const Scal&, const float* in, float* out, float*) { out[0] = in[0]; out[1] = in[0]; out[2] = in[0]; }
};

}  // namespace
