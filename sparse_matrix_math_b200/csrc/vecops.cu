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

#include "vec_functors.cuh"

namespace {

// HALO (multi-GPU CG, F = FCgPX / FCgP / FCopy3): the kernel that produces the next SpMV operand also pushes its boundary entries
// to the peers and, once every CTA is through, raises this rank's flag on them -- no separate push kernel.
template <class F, bool VEC4, bool HALO>
__global__ void __launch_bounds__(VEC_THREADS) vec_kernel(const VecParams P) {
    smm_pdl_wait();                                            // the scalars and vectors below come from the previous kernel
    bool owed = false;                                         // the solve is over but this kernel's x update is still due (FCgPX)
    if (P.state != nullptr && P.state->done) {
        if (!vec_owed<F>::value || !P.state->x_owed) return;
        owed = true;
    }
    __shared__ float red_sh[96];
    __shared__ int sh_flag;
    __shared__ HaloSeg sh_segs[HALO ? SMM_MAX_RANKS : 1];
    int nsegs = 0;
    bool pushed = false;
    if (HALO) {
        nsegs = (owed || (P.halo->debug & 4)) ? 0 : P.halo->nsegs;                      // nothing is pushed once the solve is over (on any rank: the state is the same everywhere)
        if (threadIdx.x < nsegs) sh_segs[threadIdx.x] = P.halo->segs[threadIdx.x];
        __syncthreads();
    }
    Scal sc = F::scal(P.state);
    if (vec_owed<F>::value && owed) sc.c = 1.f;
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
    if (HALO && !owed) {
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
    const int grid = smm_vec_grid(ws, a.n, aligned);
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
        case VEC_CG_R: return launch<FCgR>(a, s);
        case VEC_CG_PX: return launch<FCgPX>(a, s);
        case VEC_BICGSYM_R: return launch<FBsR>(a, s);
        case VEC_BICGSYM_PX: return launch<FBsPX>(a, s);
        default: smm_set_error("vecops: unknown kernel %d", kind); return SMM_E_INVALID;
    }
}

int smm_vec_max_grid(const smm_workspace* ws) { return ws->sm_count * VEC_CTAS_PER_SM; }

// grid of the element-wise kernels for vectors of n elements (the partial sums of their reductions depend on it: the
// persistent CG iteration reproduces the same virtual grid)
int smm_vec_grid(const smm_workspace* ws, long long n, bool aligned) {
    const long long work = aligned ? ((n + 3) >> 2) : n;
    const long long want = (work + VEC_THREADS - 1) / VEC_THREADS;
    static const int per_sm_env = [] {                         // tuning knob, read once; the workspace is sized for VEC_CTAS_PER_SM
        const char* e = getenv("SMM_B200_VEC_CTAS_PER_SM");
        const int v = e ? atoi(e) : 0;
        return v > VEC_CTAS_PER_SM ? VEC_CTAS_PER_SM : v;
    }();
    // vectors of up to two waves (L2-resident sizes) are latency-bound: half as many CTAs doing two rounds each leave
    // the last CTA half as many partial sums to fold (2 M rows: 12.3 -> 10.3 us); long vectors want every slot filled
    const int per_sm = per_sm_env > 0 ? per_sm_env : (want <= 2ll * ws->sm_count * VEC_CTAS_PER_SM ? VEC_CTAS_PER_SM / 2 : VEC_CTAS_PER_SM);
    const long long cap = (long long)ws->sm_count * per_sm;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}
