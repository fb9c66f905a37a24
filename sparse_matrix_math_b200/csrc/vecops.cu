// vecops.cu -- the per-iteration vector algebra of the Krylov solvers, fused.
//
// Each kernel is one of the reference's element-wise loops (axpy/xpay pairs) with the reduction that follows it
// in the reference folded in, and with the scalar bookkeeping (alpha/beta/omega, stopping test) executed by the
// last CTA (epilogue.cuh).  One pass over each vector per kernel; 128-bit loads/stores; a fixed grid of
// sm_count*VEC_CTAS_PER_SM CTAs with a grid-stride loop so partial sums are combined in a fixed order
// (deterministic results run to run).
//
// `_smm_fma(a,x,b)` of the reference is a*x+b with two roundings (H:27-37): smm_fma2 keeps that rounding.
#include <stdlib.h>

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

// HALO (multi-GPU CG, F = FCgP / FCopy3): the kernel that produces the next SpMV operand also pushes its boundary entries
// to the peers and, once every CTA is through, raises this rank's flag on them -- no separate push kernel.
template <class F, bool VEC4, bool HALO>
__global__ void __launch_bounds__(VEC_THREADS) vec_kernel(const VecParams P) {
    smm_pdl_wait();                                            // the scalars and vectors below come from the previous kernel
    if (P.state != nullptr && P.state->done) return;
    __shared__ float red_sh[96];
    __shared__ int sh_flag;
    __shared__ HaloSeg sh_segs[HALO ? SMM_MAX_RANKS : 1];
    int nsegs = 0;
    bool pushed = false;
    if (HALO) {
        nsegs = P.halo->nsegs;
        if (threadIdx.x < nsegs) sh_segs[threadIdx.x] = P.halo->segs[threadIdx.x];
        __syncthreads();
    }
    const Scal sc = F::scal(P.state);
    float red[2] = {0.f, 0.f};
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;

    if (VEC4) {
        const long long n4 = P.n >> 2;
        for (long long i = gid; i < n4; i += stride) {
            float4 vin[F::NIN];
#pragma unroll
            for (int k = 0; k < F::NIN; ++k) vin[k] = reinterpret_cast<const float4*>(P.in[k])[i];
            float4 vout[F::NOUT > 0 ? F::NOUT : 1];
            float ein[F::NIN], eout[F::NOUT > 0 ? F::NOUT : 1];
#define SMM_LANE(c)                                                        \
    _Pragma("unroll") for (int k = 0; k < F::NIN; ++k) ein[k] = vin[k].c;   \
    F::apply(sc, ein, eout, red);                                           \
    _Pragma("unroll") for (int k = 0; k < F::NOUT; ++k) vout[k].c = eout[k];
            SMM_LANE(x) SMM_LANE(y) SMM_LANE(z) SMM_LANE(w)
#undef SMM_LANE
#pragma unroll
            for (int k = 0; k < F::NOUT; ++k) reinterpret_cast<float4*>(P.out[k])[i] = vout[k];
            if (HALO) { const float v4[4] = {vout[0].x, vout[0].y, vout[0].z, vout[0].w}; halo_store(sh_segs, nsegs, i << 2, v4, 4, pushed); }
        }
        // tail (n % 4 elements) by the first threads of CTA 0
        const long long t = (n4 << 2) + gid;
        if (gid < 4 && t < P.n) {
            float ein[F::NIN], eout[F::NOUT > 0 ? F::NOUT : 1];
#pragma unroll
            for (int k = 0; k < F::NIN; ++k) ein[k] = P.in[k][t];
            F::apply(sc, ein, eout, red);
#pragma unroll
            for (int k = 0; k < F::NOUT; ++k) P.out[k][t] = eout[k];
            if (HALO) halo_store(sh_segs, nsegs, t, eout, 1, pushed);
        }
    } else {
        for (long long i = gid; i < P.n; i += stride) {
            float ein[F::NIN], eout[F::NOUT > 0 ? F::NOUT : 1];
#pragma unroll
            for (int k = 0; k < F::NIN; ++k) ein[k] = P.in[k][i];
            F::apply(sc, ein, eout, red);
#pragma unroll
            for (int k = 0; k < F::NOUT; ++k) P.out[k][i] = eout[k];
            if (HALO) halo_store(sh_segs, nsegs, i, eout, 1, pushed);
        }
    }

    smm_pdl_trigger();                                         // main loop done: the next kernel of the chain may start its prologue
    if (HALO) {
        if (pushed) __threadfence_system();                    // my peer stores are visible before my CTA's ticket
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(P.halo->ticket, 1u) == gridDim.x - 1) {
            __threadfence_system();
            DistComm* comm = P.halo->comm;
            const unsigned int seq = comm->push_seq + 1u;
            for (int k = 0; k < P.halo->ndests; ++k) st_release_sys_u32(comm->flags[P.halo->dests[k]] + comm->rank, seq);
            comm->push_seq = seq;
            *P.halo->ticket = 0u;
        }
    }

    if (F::NRED > 0) {
        float v[2] = {red[0], red[1]};
        if (grid_sum_last_block<2>(v, P.partials, P.partials_stride, P.ticket, red_sh, &sh_flag)) {
            if (threadIdx.x == 0) smm_finish(P.finish, P.state, v[0], v[1]);
        }
    }
}

template <class F>
int launch(const VecArgs& a, cudaStream_t s) {
    if (a.n <= 0 && F::NRED == 0) return SMM_OK;
    VecParams P;
    P.n = a.n;
    bool aligned = true;
    for (int k = 0; k < 5; ++k) { P.in[k] = a.in[k]; if (k < F::NIN && ((uintptr_t)a.in[k] & 15)) aligned = false; }
    for (int k = 0; k < 3; ++k) { P.out[k] = a.out[k]; if (k < F::NOUT && ((uintptr_t)a.out[k] & 15)) aligned = false; }
    P.state = a.state; P.finish = a.finish;
    P.partials = nullptr; P.partials_stride = 0; P.ticket = nullptr;
    P.halo = static_cast<const HaloPushDev*>(a.halo_push);
    if (P.halo != nullptr && !(F::NOUT >= 1 && F::NRED == 0 && F::HALO_OK)) { smm_set_error("vecops: this kernel cannot push a halo"); return SMM_E_INVALID; }
    smm_workspace* ws = a.ws;
    long long work = aligned ? ((a.n + 3) >> 2) : a.n;
    long long want = (work + VEC_THREADS - 1) / VEC_THREADS;
    static const int per_sm_env = [] {                         // tuning knob, read once; the workspace is sized for VEC_CTAS_PER_SM
        const char* e = getenv("SMM_B200_VEC_CTAS_PER_SM");
        const int v = e ? atoi(e) : 0;
        return v > VEC_CTAS_PER_SM ? VEC_CTAS_PER_SM : v;
    }();
    // vectors of up to two waves (L2-resident sizes) are latency-bound: half as many CTAs doing two rounds each leave
    // the last CTA half as many partial sums to fold (2 M rows: 12.3 -> 10.3 us); long vectors want every slot filled
    const int per_sm = per_sm_env > 0 ? per_sm_env : (want <= 2ll * ws->sm_count * VEC_CTAS_PER_SM ? VEC_CTAS_PER_SM / 2 : VEC_CTAS_PER_SM);
    const long long cap = (long long)ws->sm_count * per_sm;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    if (F::NRED > 0) {
        if (ws->partials_cap < (size_t)grid) { smm_set_error("vecops: reduction workspace too small"); return SMM_E_STATE; }
        P.partials = ws->partials + (size_t)a.slot * 2 * ws->partials_cap;
        P.partials_stride = ws->partials_cap;
        P.ticket = ws->tickets + a.slot;
    }
    cudaError_t le;
    if (F::HALO_OK && P.halo != nullptr) {
        if (aligned) le = smm_launch_chain(vec_kernel<F, true, F::HALO_OK>, grid, VEC_THREADS, 0, s, P);
        else le = smm_launch_chain(vec_kernel<F, false, F::HALO_OK>, grid, VEC_THREADS, 0, s, P);
    } else if (aligned) le = smm_launch_chain(vec_kernel<F, true, false>, grid, VEC_THREADS, 0, s, P);
    else le = smm_launch_chain(vec_kernel<F, false, false>, grid, VEC_THREADS, 0, s, P);
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(le);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

}  // namespace

int smm_launch_vec(int kind, const VecArgs& a, cudaStream_t s) {
    switch (kind) {
        case VEC_CG_XR: return launch<FCgXR>(a, s);
        case VEC_CG_P: return launch<FCgP>(a, s);
        case VEC_BICGSYM_XR: return launch<FBsXR>(a, s);
        case VEC_BICGSYM_P: return launch<FBsP>(a, s);
        case VEC_CGS_QX: return launch<FCgsQX>(a, s);
        case VEC_CGS_UP: return launch<FCgsUP>(a, s);
        case VEC_STAB_S: return launch<FStabS>(a, s);
        case VEC_STAB_XR: return launch<FStabXR>(a, s);
        case VEC_STAB_P: return launch<FStabP>(a, s);
        case VEC_DOT2: return launch<FDot2>(a, s);
        case VEC_COPY3: return launch<FCopy3>(a, s);
        default: smm_set_error("vecops: unknown kernel %d", kind); return SMM_E_INVALID;
    }
}

int smm_vec_max_grid(const smm_workspace* ws) { return ws->sm_count * VEC_CTAS_PER_SM; }
