// solvers.cu -- iteration drivers of the four Krylov solvers.
//
//   smm_solve_cg        <- SMM::ConjugateGradient        H:2316-2398
//   smm_solve_bicgsym   <- SMM::BiCGSymmetric            H:2021-2102
//   smm_solve_cgs       <- SMM::ConjugateGradientSquared H:2109-2178
//   smm_solve_bicgstab  <- SMM::BiCGStab                 H:2191-2303
//
// Every scalar of the recurrences (alpha, beta, omega, the residual norms), the iteration counter and the
// stopping / DIVERGED tests live in a SolveState block in HBM and are updated by the last CTA of the kernel that
// produced the reduction (epilogue.cuh), so the host never reads a dot product.  One iteration is a fixed
// sequence of kernels; every kernel starts with `if (state->done) return`, which makes the iterations enqueued
// after convergence free of side effects.  The loop itself is
//   GRAPH_WHILE    a CUDA-graph conditional WHILE node: one launch, the device decides when to stop;
//   GRAPH_CHUNKED  the captured iteration graph launched `check_every` times between polls of the done flag;
//   STREAM         plain launches (debug).
//
// Fused kernel sequence per iteration (FAST reductions) and algorithmic bytes (n rows, nnz entries):
//   CG        SpMV[Ap=A p; p.Ap->alpha] | x,r update + r.r -> beta, stop | p update            8 nnz + 48 n
//   BiCGSym   same shape (different rounding and tests)                                         8 nnz + 48 n
//   CGS       SpMV[ap; ap.r0->alpha] | q,auq,x | SpMV[r-=A auq; r.r0, r.r->beta, stop] | u,p    16 nnz + 80 n
//   BiCGStab  SpMV[ap; ap.r0->alpha] | s | SpMV[as; as.s, as.as->omega] | x,r + r.r, r.r0 | p   16 nnz + 84 n
// In the REFERENCE_* reduction modes the reductions are un-fused and use dots.cu so that every float operation
// happens in the reference's order.
#include <chrono>
#include <stdlib.h>
#include <string.h>

#include "dist.h"
#include "epilogue.cuh"
#include "smm_internal.cuh"

// AUTO takes the persistent CG iteration for L2-resident problems when this is 1 (set from the measurement in
// profiles/r02_driver_bench.txt)
#ifndef SMM_PERSISTENT_AUTO
#define SMM_PERSISTENT_AUTO 0
#endif

namespace {

struct Ctx {
    smm_dist* dist = nullptr;        // multi-GPU: rows of this rank, extended vector, peer mailboxes

    const smm_csr* a = nullptr;
    const smm_precond* precond = nullptr;
    smm_workspace* ws = nullptr;
    cudaStream_t s = nullptr;
    SolveState* st = nullptr;
    long long n = 0;
    int mode = SMM_REDUCE_FAST;      // reduction mode
    bool exact = false;              // mode != FAST
    const float* b = nullptr;
    float* x = nullptr;
    float *r = nullptr, *p = nullptr, *ap = nullptr, *r0 = nullptr, *u = nullptr, *q = nullptr, *auq = nullptr,
          *sv = nullptr, *as = nullptr, *scratch = nullptr;
    int kernels_per_iteration = 0;
};

// multi-GPU: can the SpMV wait for the halo itself (rows kernel), or does it need the wait kernel in front of it?
// Default since the end of round 2: separate push and wait kernels.  Measured with the final kernels (profiles/r02_dist_overhead.txt
// section 5): 2 GPUs 1042.9 us per iteration against 1114.4 (halo stores fused into the x,p update + wait kernel) and 1086.8 (fused
// stores + wait inside the SpMV); 8 GPUs 309.3 against 308.4 -- never slower, and the x,p update and the SpMV stay the plain kernels
// of the single-GPU solve.  SMM_B200_DIST_FUSED=1 selects the fused forms.
bool dist_fused() { static const bool on = [] { const char* e = getenv("SMM_B200_DIST_FUSED"); return e && atoi(e) != 0; }(); return on; }
// With the fused stores (SMM_B200_DIST_FUSED=1): the wait for the peers' flags is a one-warp kernel in front of the plain SpMV, or
// (SMM_B200_DIST_FUSED_WAIT=1) happens inside the SpMV's HALO instance, which walks the interior row groups first -- that
// instance is 3 % slower than the plain one on the same rows (profiles/r02_dist_overhead.txt).
bool dist_fused_wait_on() { static const bool on = [] { const char* e = getenv("SMM_B200_DIST_FUSED_WAIT"); return e && atoi(e) != 0; }(); return on; }
bool fused_wait(const Ctx& c) { return dist_fused() && dist_fused_wait_on() && c.dist && c.dist->nranks > 1 && c.dist->wait_dev && smm_spmv_rows_lanes(c.a, c.exact ? 1 : 0) > 0; }

int spmv(Ctx& c, int op, const float* lhs, const float* mult, float* out, int reduce, int finish, const float* aux,
         float* c1 = nullptr, float* c2 = nullptr, float* c3 = nullptr) {
    SpmvArgs a;
    if (c.dist) {
        smm_dist* d = c.dist;
        if (mult != d->ext) {
            // the operand's owned entries go into the extended vector, the halo comes from the peers
            if (mult != d->ext + d->own_off)
                SMM_CUDA(cudaMemcpyAsync(d->ext + d->own_off, mult, sizeof(float) * (size_t)c.n, cudaMemcpyDeviceToDevice, c.s));
            SMM_TRY(smm_dist_exchange_async(d, c.st, c.s, !fused_wait(c)));
            mult = d->ext;
        }   // else: the kernel that produced the operand has pushed its boundary already (CG's p)
        if (fused_wait(c)) a.halo_wait = d->wait_dev;
    }
    a.m = c.a; a.op = op; a.lhs = lhs; a.mult = mult; a.out = out; a.exact = c.exact ? 1 : 0;
    a.reduce = reduce; a.finish = finish; a.slot = 0; a.aux = aux; a.state = c.st;
    a.copy1 = c1; a.copy2 = c2; a.copy3 = c3;
    return smm_launch_spmv(a, c.s);
}

int vec(Ctx& c, int kind, int finish, std::initializer_list<const float*> in, std::initializer_list<float*> out, bool push_halo = false) {
    VecArgs v;
    v.n = c.n; v.state = c.st; v.finish = finish; v.slot = 1; v.ws = c.ws;
    if (push_halo && c.dist && c.dist->nranks > 1 && dist_fused()) v.halo_push = c.dist->push_dev;
    int i = 0;
    for (const float* p : in) v.in[i++] = p;
    i = 0;
    for (float* p : out) v.out[i++] = p;
    return smm_launch_vec(kind, v, c.s);
}

// a dot product (or two) whose totals feed `finish`; FAST: fused two-stage kernel, otherwise the reference order
int dots(Ctx& c, int finish, const float* a0, const float* b0, const float* a1 = nullptr, const float* b1 = nullptr) {
    if (c.mode == SMM_REDUCE_FAST) {
        // VEC_DOT2 computes (a.b, a.a): callers in FAST mode only use shapes it covers
        return vec(c, VEC_DOT2, finish, {a0, b0}, {});
    }
    return smm_launch_dot_ref(c.mode, c.n, a1 ? 2 : 1, a0, b0, a1 ? a1 : a0, b1 ? b1 : b0, c.st, finish, nullptr, c.s, c.ws);
}

// ---------------------------------------------------------------------------------------------------
// per-solver: start-up sequence and one iteration
// ---------------------------------------------------------------------------------------------------
// Multi-GPU CG: p lives inside the extended vector (owned part + halo); every SpMV operand is exchanged first.
// The reductions are summed over the ranks inside the kernels' epilogues (dist_device.cuh), so the scalar state --
// and with it every branch -- is bit-identical on all ranks.
// p lives in the extended vector.  Default: after the kernel that writes it (p = r at the start, the x,p update afterwards) a push
// kernel stores its boundary entries into the peers' extended vectors and raises the flags, a wait kernel spins on the peers'
// flags, and the SpMV is the plain kernel.  Fused forms (SMM_B200_DIST_FUSED=1): the writing kernel stores the boundary entries
// itself, and (SMM_B200_DIST_FUSED_WAIT=1) the SpMV multiplies its interior rows while those stores travel, waiting for the
// peers' flags only in the warps that reach a boundary row group.
// after the kernel that wrote p: whatever part of the exchange that kernel and the next SpMV do not do themselves
int dist_after_p(Ctx& c) {
    if (!dist_fused()) return smm_dist_exchange_async(c.dist, c.st, c.s, true);
    if (!fused_wait(c)) return smm_dist_wait_async(c.dist, c.st, c.s);
    return SMM_OK;
}
int cg_init_dist(Ctx& c, const float* x0) {
    smm_dist* d = c.dist;
    if (c.exact) {                                             // reference-tree mode: local tree, ranks joined pairwise
        SMM_TRY(spmv(c, SMM_OP_SUB, c.b, x0, c.r, RED_NONE, FIN_NONE, nullptr));
        SMM_TRY(dots(c, FIN_CG_INIT, c.r, c.r));
    } else
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, x0, c.r, RED_OUT_OUT, FIN_CG_INIT, nullptr));
    SMM_TRY(vec(c, VEC_COPY3, FIN_NONE, {c.r}, {c.p, c.p, c.p}, true));
    return dist_after_p(c);
}
int cg_iter_dist(Ctx& c) {
    smm_dist* d = c.dist;
    const int extra = !dist_fused() ? 2 : (fused_wait(c) ? 0 : 1);
    if (c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, d->ext, c.ap, RED_NONE, FIN_NONE, nullptr));
        SMM_TRY(dots(c, FIN_CG_ALPHA, c.ap, c.p));
        const bool fused_r = smm_dot_ref_update_applies(c.mode, c.n, c.r, c.ap);   // the r update rides on the tree dot that follows it
        if (fused_r) SMM_TRY(smm_launch_dot_ref(c.mode, c.n, 1, c.r, c.ap, c.r, c.ap, c.st, FIN_CG_UPDATE, nullptr, c.s, c.ws, c.r));
        else {
            SMM_TRY(vec(c, VEC_CG_R, FIN_NONE, {c.r, c.ap}, {c.r}));
            SMM_TRY(dots(c, FIN_CG_UPDATE, c.r, c.r));
        }
        SMM_TRY(vec(c, VEC_CG_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}, true));
        SMM_TRY(dist_after_p(c));
        c.kernels_per_iteration = (fused_r ? 4 : 5) + extra;
        return SMM_OK;
    }
    SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, d->ext, c.ap, RED_OUT_AUX, FIN_CG_ALPHA, c.p));
    SMM_TRY(vec(c, VEC_CG_R, FIN_CG_UPDATE, {c.r, c.ap}, {c.r}));
    SMM_TRY(vec(c, VEC_CG_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}, true));
    SMM_TRY(dist_after_p(c));
    c.kernels_per_iteration = 3 + extra;
    return SMM_OK;
}

int cg_init(Ctx& c, const float* x0) {
    // r = b - A x0 ; p = r ; rr = r.r ; early SUCCESS if eps^2 > rr          H:2336-2347
    if (!c.exact) return spmv(c, SMM_OP_SUB, c.b, x0, c.r, RED_OUT_OUT, FIN_CG_INIT, nullptr, c.p);
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, x0, c.r, RED_NONE, FIN_NONE, nullptr, c.p));
    return dots(c, FIN_CG_INIT, c.r, c.r);
}
int cg_iter(Ctx& c) {
    if (!c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_OUT_AUX, FIN_CG_ALPHA, c.p));        // H:2353-2358
        // two passes that read p once: r and the stopping test first, then x (with this iteration's p) and the new p
        SMM_TRY(vec(c, VEC_CG_R, FIN_CG_UPDATE, {c.r, c.ap}, {c.r}));                              // H:2366-2382
        SMM_TRY(vec(c, VEC_CG_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}));                         // H:2363-2365, H:2385-2393
        c.kernels_per_iteration = 3;
        return SMM_OK;
    }
    SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_NONE, FIN_NONE, nullptr));
    SMM_TRY(dots(c, FIN_CG_ALPHA, c.ap, c.p));
    const bool fused_r = smm_dot_ref_update_applies(c.mode, c.n, c.r, c.ap);       // the r update rides on the tree dot that follows it
    if (fused_r) SMM_TRY(smm_launch_dot_ref(c.mode, c.n, 1, c.r, c.ap, c.r, c.ap, c.st, FIN_CG_UPDATE, nullptr, c.s, c.ws, c.r));
    else {
        SMM_TRY(vec(c, VEC_CG_R, FIN_NONE, {c.r, c.ap}, {c.r}));
        SMM_TRY(dots(c, FIN_CG_UPDATE, c.r, c.r));
    }
    SMM_TRY(vec(c, VEC_CG_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}));
    c.kernels_per_iteration = fused_r ? 4 : 5;
    return SMM_OK;
}

// ConjugateGradient with the IC(0) preconditioner, H:2414-2505.  z lives in c.sv.
int pcg_init(Ctx& c, const float* x0) {
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, x0, c.r, RED_NONE, FIN_NONE, nullptr));                        // H:2440
    SMM_TRY(smm_sgs_apply_async(c.precond, c.r, c.sv, c.st, c.s));                                 // z = M^-1 r, H:2441
    SMM_TRY(vec(c, VEC_COPY3, FIN_NONE, {c.sv}, {c.p, c.p, c.p}));                                 // p = z, H:2447
    // rz and ||r||^2 are accumulated by a plain loop in BOTH builds of the reference (H:2444-2448)
    if (c.mode == SMM_REDUCE_FAST) return vec(c, VEC_DOT2, FIN_PCG_INIT, {c.r, c.sv}, {});
    return smm_launch_dot_ref(SMM_REDUCE_REFERENCE_SERIAL, c.n, 2, c.r, c.sv, c.r, c.r, c.st, FIN_PCG_INIT, nullptr, c.s, c.ws);
}
int pcg_iter(Ctx& c) {
    if (!c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_OUT_AUX, FIN_PCG_ALPHA, c.p));       // H:2461-2466
    } else {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_NONE, FIN_NONE, nullptr));
        SMM_TRY(dots(c, FIN_PCG_ALPHA, c.ap, c.p));
    }
    SMM_TRY(vec(c, VEC_CG_XR, FIN_NONE, {c.x, c.p, c.r, c.ap}, {c.x, c.r}));                       // H:2470-2480
    SMM_TRY(smm_sgs_apply_async(c.precond, c.r, c.sv, c.st, c.s));                                 // H:2482
    if (!c.exact) SMM_TRY(vec(c, VEC_DOT2, FIN_PCG_UPDATE, {c.r, c.sv}, {}));                      // (r.z, r.r), H:2483-2488
    else SMM_TRY(dots(c, FIN_PCG_UPDATE, c.r, c.sv, c.r, c.r));
    SMM_TRY(vec(c, VEC_CG_P, FIN_NONE, {c.p, c.sv}, {c.p}));                                       // p = fma(beta, p, z), H:2490-2499
    c.kernels_per_iteration = (c.exact ? 5 : 4) + smm_sgs_kernels_per_apply(c.precond);
    return SMM_OK;
}

int bicgsym_init(Ctx& c) {
    if (!c.exact) return spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_OUT_OUT, FIN_RR_INIT, nullptr, c.p);   // H:2035-2043
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_NONE, FIN_NONE, nullptr, c.p));
    return dots(c, FIN_RR_INIT, c.r, c.r);
}
int bicgsym_iter(Ctx& c) {
    if (!c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_OUT_AUX, FIN_BICGSYM_ALPHA, c.p));   // H:2048-2059
        SMM_TRY(vec(c, VEC_BICGSYM_R, FIN_BICGSYM_UPDATE, {c.r, c.ap}, {c.r}));                    // H:2068-2082, 2094-2096
        SMM_TRY(vec(c, VEC_BICGSYM_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}));                    // H:2061-2067, H:2084-2092 (p read once)
        c.kernels_per_iteration = 3;
        return SMM_OK;
    }
    SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_NONE, FIN_NONE, nullptr));
    SMM_TRY(dots(c, FIN_BICGSYM_ALPHA, c.ap, c.p));
    SMM_TRY(vec(c, VEC_BICGSYM_R, FIN_NONE, {c.r, c.ap}, {c.r}));
    SMM_TRY(dots(c, FIN_BICGSYM_UPDATE, c.r, c.r));
    SMM_TRY(vec(c, VEC_BICGSYM_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}));
    c.kernels_per_iteration = 5;
    return SMM_OK;
}

int cgs_init(Ctx& c) {
    // r = b - A x ; p = u = r0 = r ; rr0 = r.r0                                 H:2117-2128
    if (!c.exact) return spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_OUT_OUT, FIN_RR_INIT, nullptr, c.p, c.u, c.r0);
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_NONE, FIN_NONE, nullptr, c.p, c.u, c.r0));
    return dots(c, FIN_RR_INIT, c.r, c.r0);
}
int cgs_iter(Ctx& c) {
    if (!c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_OUT_AUX, FIN_ALPHA_R0, c.r0));       // H:2132-2135
        SMM_TRY(vec(c, VEC_CGS_QX, FIN_NONE, {c.ap, c.u, c.x}, {c.q, c.auq, c.x}));                // H:2137-2149
        SMM_TRY(spmv(c, SMM_OP_SUB, c.r, c.auq, c.r, RED_OUT_AUX_OUT_OUT, FIN_CGS_UPDATE, c.r0));   // H:2151-2154, 2169-2172
        SMM_TRY(vec(c, VEC_CGS_UP, FIN_NONE, {c.q, c.r, c.p}, {c.u, c.p}));                        // H:2157-2167
        c.kernels_per_iteration = 4;
        return SMM_OK;
    }
    SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_NONE, FIN_NONE, nullptr));
    SMM_TRY(dots(c, FIN_ALPHA_R0, c.ap, c.r0));
    SMM_TRY(vec(c, VEC_CGS_QX, FIN_NONE, {c.ap, c.u, c.x}, {c.q, c.auq, c.x}));
    SMM_TRY(spmv(c, SMM_OP_SUB, c.r, c.auq, c.r, RED_NONE, FIN_NONE, nullptr));
    SMM_TRY(dots(c, FIN_CGS_UPDATE, c.r, c.r0, c.r, c.r));
    SMM_TRY(vec(c, VEC_CGS_UP, FIN_NONE, {c.q, c.r, c.p}, {c.u, c.p}));
    c.kernels_per_iteration = 6;
    return SMM_OK;
}

int stab_init(Ctx& c) {
    // r = b - A x ; [r = M^-1 r] ; r0 = p = r ; rr0 = r.r0                       H:2214-2231
    if (!c.precond) {
        if (!c.exact) return spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_OUT_OUT, FIN_RR_INIT, nullptr, c.p, c.r0);
        SMM_TRY(spmv(c, SMM_OP_SUB, c.b, c.x, c.r, RED_NONE, FIN_NONE, nullptr, c.p, c.r0));
        return dots(c, FIN_RR_INIT, c.r, c.r0);
    }
    SMM_TRY(spmv(c, SMM_OP_SUB, c.b, c.x, c.scratch, RED_NONE, FIN_NONE, nullptr));
    SMM_TRY(smm_sgs_apply_async(c.precond, c.scratch, c.r, c.st, c.s));
    SMM_TRY(vec(c, VEC_COPY3, FIN_NONE, {c.r}, {c.r0, c.p, c.r0}));
    return dots(c, FIN_RR_INIT, c.r, c.r0);
}
int stab_iter(Ctx& c) {
    int k = 0;
    if (!c.precond && !c.exact) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_OUT_AUX, FIN_ALPHA_R0, c.r0));       // H:2240-2244
        SMM_TRY(vec(c, VEC_STAB_S, FIN_NONE, {c.ap, c.r}, {c.sv}));                                // H:2245-2247
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.sv, c.as, RED_OUT_AUX_OUT_OUT, FIN_BICGSTAB_OMEGA, c.sv));   // H:2256-2261
        SMM_TRY(vec(c, VEC_STAB_XR, FIN_BICGSTAB_UPDATE, {c.x, c.p, c.sv, c.as, c.r0}, {c.x, c.r})); // H:2262-2271, 2275-2277
        SMM_TRY(vec(c, VEC_STAB_P, FIN_NONE, {c.p, c.ap, c.r}, {c.p}));                            // H:2272-2274
        c.kernels_per_iteration = 5;
        return SMM_OK;
    }
    // ap = [M^-1] A p                                                             H:2233-2241
    if (c.precond) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.scratch, RED_NONE, FIN_NONE, nullptr));
        SMM_TRY(smm_sgs_apply_async(c.precond, c.scratch, c.ap, c.st, c.s));
        k += 1 + smm_sgs_kernels_per_apply(c.precond);
    } else {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.p, c.ap, RED_NONE, FIN_NONE, nullptr));
        k += 1;
    }
    SMM_TRY(dots(c, FIN_ALPHA_R0, c.ap, c.r0));
    SMM_TRY(vec(c, VEC_STAB_S, FIN_NONE, {c.ap, c.r}, {c.sv}));
    // as = [M^-1] A s                                                             H:2249-2257
    if (c.precond) {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.sv, c.scratch, RED_NONE, FIN_NONE, nullptr));
        SMM_TRY(smm_sgs_apply_async(c.precond, c.scratch, c.as, c.st, c.s));
        k += 1 + smm_sgs_kernels_per_apply(c.precond);
    } else {
        SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, c.sv, c.as, RED_NONE, FIN_NONE, nullptr));
        k += 1;
    }
    if (c.mode == SMM_REDUCE_FAST) {
        SMM_TRY(dots(c, FIN_BICGSTAB_OMEGA, c.as, c.sv));                                          // (as.s, as.as)
        SMM_TRY(vec(c, VEC_STAB_XR, FIN_BICGSTAB_UPDATE, {c.x, c.p, c.sv, c.as, c.r0}, {c.x, c.r}));
        k += 4;
    } else {
        SMM_TRY(dots(c, FIN_BICGSTAB_OMEGA, c.as, c.sv, c.as, c.as));
        SMM_TRY(vec(c, VEC_STAB_XR, FIN_NONE, {c.x, c.p, c.sv, c.as, c.r0}, {c.x, c.r}));
        // resL2Norm is a serial left-to-right sum in BOTH builds of the reference (H:2262-2267); r.r0 follows the build
        SMM_TRY(smm_launch_dot_ref(SMM_REDUCE_REFERENCE_SERIAL, c.n, 1, c.r, c.r, c.r, c.r, c.st, FIN_STASH0, nullptr, c.s, c.ws));
        SMM_TRY(dots(c, FIN_BICGSTAB_UPDATE_STASHED, c.r, c.r0));
        k += 6;
    }
    SMM_TRY(vec(c, VEC_STAB_P, FIN_NONE, {c.p, c.ap, c.r}, {c.p}));
    c.kernels_per_iteration = k + 1;
    return SMM_OK;
}

// ---------------------------------------------------------------------------------------------------
// loop drivers
// ---------------------------------------------------------------------------------------------------
__global__ void while_condition_kernel(const SolveState* st, cudaGraphConditionalHandle h) {
    cudaGraphSetConditional(h, st->done ? 0u : 1u);
}

typedef int (*IterFn)(Ctx&);

int poll_state(Ctx& c) {
    SMM_CUDA(cudaMemcpyAsync(c.ws->state_host, c.st, sizeof(SolveState), cudaMemcpyDeviceToHost, c.s));
    SMM_CUDA(cudaStreamSynchronize(c.s));
    return SMM_OK;
}

int run_stream(Ctx& c, IterFn iter, long long budget, int check_every, long long* launches) {
    long long done_iters = 0;
    while (done_iters < budget) {
        const long long m = (budget - done_iters < check_every) ? budget - done_iters : check_every;
        for (long long i = 0; i < m; ++i) SMM_TRY(iter(c));
        *launches += m * c.kernels_per_iteration;
        done_iters += m;
        SMM_TRY(poll_state(c));
        if (c.ws->state_host->done) break;
    }
    return SMM_OK;
}

int capture_iteration(Ctx& c, IterFn iter, cudaGraph_t* graph) {
    SMM_CUDA(cudaStreamBeginCapture(c.s, cudaStreamCaptureModeThreadLocal));
    t_smm_capturing = true;                                   // captured, not launched: counted per graph launch
    int rc = iter(c);
    t_smm_capturing = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c.s, &g);
    if (rc != SMM_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return smm_cuda_fail(e, "cudaStreamEndCapture", __FILE__, __LINE__);
    *graph = g;
    return SMM_OK;
}

int run_graph_chunked(Ctx& c, IterFn iter, long long budget, int check_every, long long* launches) {
    cudaGraph_t g = nullptr;
    SMM_TRY(capture_iteration(c, iter, &g));
    cudaGraphExec_t exec = nullptr;
    cudaError_t e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return smm_cuda_fail(e, "cudaGraphInstantiate", __FILE__, __LINE__); }
    int rc = SMM_OK;
    long long done_iters = 0;
    while (done_iters < budget && rc == SMM_OK) {
        const long long m = (budget - done_iters < check_every) ? budget - done_iters : check_every;
        for (long long i = 0; i < m && rc == SMM_OK; ++i) {
            e = cudaGraphLaunch(exec, c.s);
            if (e != cudaSuccess) rc = smm_cuda_fail(e, "cudaGraphLaunch", __FILE__, __LINE__);
        }
        *launches += m * c.kernels_per_iteration;
        g_smm_launches.fetch_add(m * c.kernels_per_iteration, std::memory_order_relaxed);
        done_iters += m;
        if (rc == SMM_OK) rc = poll_state(c);
        if (rc == SMM_OK && c.ws->state_host->done) break;
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(g);
    return rc;
}

int run_graph_while(Ctx& c, IterFn iter, long long* launches) {
    cudaGraph_t g = nullptr;
    SMM_CUDA(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h;
    cudaError_t e = cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return smm_cuda_fail(e, "cudaGraphConditionalHandleCreate", __FILE__, __LINE__); }
    cudaGraphNodeParams np = {cudaGraphNodeTypeConditional};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = h;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    e = cudaGraphAddNode(&node, g, nullptr, 0, &np);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return smm_cuda_fail(e, "cudaGraphAddNode(conditional)", __FILE__, __LINE__); }
    cudaGraph_t body = np.conditional.phGraph_out[0];
    e = cudaStreamBeginCaptureToGraph(c.s, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return smm_cuda_fail(e, "cudaStreamBeginCaptureToGraph", __FILE__, __LINE__); }
    t_smm_capturing = true;
    int rc = iter(c);
    if (rc == SMM_OK) {
        while_condition_kernel<<<1, 1, 0, c.s>>>(c.st, h);
        if (cudaGetLastError() != cudaSuccess) rc = SMM_E_CUDA;
    }
    t_smm_capturing = false;
    e = cudaStreamEndCapture(c.s, nullptr);
    if (rc != SMM_OK || e != cudaSuccess) {
        cudaGraphDestroy(g);
        return rc != SMM_OK ? rc : smm_cuda_fail(e, "cudaStreamEndCapture(body)", __FILE__, __LINE__);
    }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return smm_cuda_fail(e, "cudaGraphInstantiate(while)", __FILE__, __LINE__); }
    e = cudaGraphLaunch(exec, c.s);
    if (e != cudaSuccess) rc = smm_cuda_fail(e, "cudaGraphLaunch(while)", __FILE__, __LINE__);
    if (rc == SMM_OK) rc = poll_state(c);
    if (rc == SMM_OK) {
        // the body ran once per iteration plus possibly one no-op pass
        const long long it = c.ws->state_host->iterations;
        *launches += (it + 1) * (c.kernels_per_iteration + 1);
        g_smm_launches.fetch_add((it + 1) * (c.kernels_per_iteration + 1), std::memory_order_relaxed);
    }
    cudaGraphExecDestroy(exec);
    cudaGraphDestroy(g);
    return rc;
}

// ---------------------------------------------------------------------------------------------------
// common front end
// ---------------------------------------------------------------------------------------------------
enum Solver { S_CG, S_BICGSYM, S_CGS, S_BICGSTAB, S_CG_IC0 };

int clamp_iterations(int solver, int max_iterations, int rows) {
    if (solver == S_CG || solver == S_CG_IC0) return max_iterations == -1 ? rows : max_iterations;   // H:2345-2347, H:2452-2454
    int m = max_iterations < rows ? max_iterations : rows;                                          // H:2030, 2111, 2200
    if (m == -1) m = rows;                                                                          // H:2031-2033
    return m;
}

int solve_dev(int solver, const smm_csr* a, const smm_precond* precond, const float* b_dev, const float* x0_dev, float* x_dev,
              int max_iterations, float eps, const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s,
              smm_dist* dist = nullptr) {
    if (!a || (a->rows && (!b_dev || !x_dev || !x0_dev))) { smm_set_error("solve: bad arguments"); return SMM_E_INVALID; }
    if (!dist && a->rows != a->cols) { smm_set_error("solve: matrix must be square"); return SMM_E_INVALID; }
    SMM_CUDA(cudaSetDevice(a->device));
    smm_workspace* ws = nullptr;
    SMM_TRY(smm_workspace_get(a, &ws));
    Ctx c;
    c.a = a; c.precond = precond; c.ws = ws; c.s = s; c.st = ws->state; c.n = a->rows; c.dist = dist;
    c.mode = opts ? opts->reduction_mode : SMM_REDUCE_FAST;
    if (c.mode < 0 || c.mode > 2) { smm_set_error("solve: unknown reduction mode"); return SMM_E_INVALID; }
    if (dist && c.mode != SMM_REDUCE_FAST) {
        // The reference-tree mode distributes when every rank owns one node of the reference's reduction tree (the range
        // [0, n) halved log2(P) times at lo + (hi - lo) / 2, H:308-320): local trees, then the ranks joined pairwise.
        // (BiCGStab's ||r||^2 is one serial sum over the whole vector in the reference, H:2262-2267: it is chained through the
        // ranks, dots.cu / dist_device.cuh.)
        bool aligned = c.mode == SMM_REDUCE_REFERENCE_TREE && (dist->nranks & (dist->nranks - 1)) == 0;
        if (aligned) {
            long long lo = 0, hi = dist->global_rows;
            for (int bit = dist->nranks >> 1; bit >= 1; bit >>= 1) {
                const long long mid = lo + (hi - lo) / 2;
                if (dist->rank & bit) lo = mid; else hi = mid;
            }
            aligned = lo == dist->row_begin && hi == dist->row_end && hi - lo > 8192;
        }
        if (!aligned) {
            smm_set_error("multi-GPU solve: of the reference-order modes only REFERENCE_TREE distributes, with a power-of-two "
                          "number of ranks and row blocks that are the nodes of the reference's reduction tree (dist.tbb_partition)");
            return SMM_E_INVALID;
        }
    }
    if (dist) {
        const int tree_order = c.mode == SMM_REDUCE_REFERENCE_TREE ? 1 : 0;
        SMM_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(dist->comm_dev) + offsetof(DistComm, tree_order), &tree_order, sizeof(int), cudaMemcpyHostToDevice, s));
        SMM_CUDA(cudaStreamSynchronize(s));                    // the source is a stack variable
    }
    c.exact = c.mode != SMM_REDUCE_FAST;
    if (c.exact) SMM_TRY(smm_dot_ref_prepare(a->rows, ws));        // scratch must exist before the iteration graph is captured
    c.b = b_dev; c.x = x_dev;
    const int nvec = solver == S_CG || solver == S_BICGSYM ? 3 : (solver == S_CGS ? 7 : 7);
    // work vectors live behind the three host-I/O staging slots (vec[0..2])
    SMM_TRY(smm_workspace_vectors(ws, 3 + nvec, (size_t)a->rows));
    float** w = ws->vec + 3;
    c.r = w[0]; c.p = w[1]; c.ap = w[2];
    if (dist && solver == S_CG) c.p = dist->ext + dist->own_off;   // CG keeps p inside the extended vector (no staging copy)
    if (solver == S_CGS) { c.r0 = w[3]; c.u = w[4]; c.q = w[5]; c.auq = w[6]; }
    if (solver == S_BICGSTAB) { c.r0 = w[3]; c.sv = w[4]; c.as = w[5]; c.scratch = w[6]; }
    if (solver == S_CG_IC0) { c.sv = w[3]; }

    const int driver_req = opts ? opts->driver_mode : SMM_DRIVER_AUTO;
    if (driver_req < SMM_DRIVER_AUTO || driver_req > SMM_DRIVER_PERSISTENT) { smm_set_error("solve: unknown driver mode"); return SMM_E_INVALID; }
    int driver = driver_req == SMM_DRIVER_AUTO ? SMM_DRIVER_GRAPH_CHUNKED : driver_req;
    // the persistent iteration exists for ConjugateGradient with fast reductions on one GPU; AUTO takes it when the whole
    // working set (matrix + the four vectors of the loop) lives in L2 (SMM_B200_PERSISTENT=0 / 1 forces the choice for AUTO)
    const bool persistent_ok = solver == S_CG && !dist && !c.exact && a->rows > 0;
    if (driver == SMM_DRIVER_PERSISTENT && !persistent_ok) driver = SMM_DRIVER_GRAPH_CHUNKED;
    if (driver_req == SMM_DRIVER_AUTO && persistent_ok) {
        static const int forced = [] { const char* e = getenv("SMM_B200_PERSISTENT"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
        int l2 = 0;
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, a->device);
        const double working = 8.0 * (double)a->nnz + 4.0 * ((double)a->rows + 1) + 16.0 * (double)a->rows;
        if (forced == 1 || (forced < 0 && SMM_PERSISTENT_AUTO && working <= 0.75 * (double)l2)) driver = SMM_DRIVER_PERSISTENT;
    }
    int check_every = opts && opts->check_every > 0 ? opts->check_every : 32;
    const int hist_cap = opts && opts->history && opts->history_cap > 0 ? opts->history_cap : 0;
    if (hist_cap > ws->history_cap) {
        cudaFree(ws->history);
        ws->history = nullptr;
        SMM_CUDA(cudaMalloc(&ws->history, sizeof(float) * (size_t)hist_cap));
        ws->history_cap = hist_cap;
    }

    SolveState* h = ws->state_host;
    memset(h, 0, sizeof *h);
    h->max_iterations = clamp_iterations(solver, max_iterations, a->rows);
    h->eps = eps;
    h->eps2 = eps * eps;                                       // H:2045, H:2130, H:2335 (float product)
    h->history = hist_cap ? ws->history : nullptr;
    h->history_cap = hist_cap;
    h->comm = (dist && dist->nranks > 1) ? dist->comm_dev : nullptr;
    // the iteration cap refers to the GLOBAL system (H:2345-2347: -1 means rows)
    if (dist) h->max_iterations = clamp_iterations(solver, max_iterations, (int)dist->global_rows);
    const int max_it = h->max_iterations;
    SMM_CUDA(cudaMemcpyAsync(c.st, h, sizeof *h, cudaMemcpyHostToDevice, s));

    long long launches = 0;
    const long long launches_before = t_smm_launches;
    SMM_CUDA(cudaEventRecord(ws->ev0, s));
    IterFn iter = nullptr;
    long long budget = 0;
    if (a->rows > 0) {
        switch (solver) {
            case S_CG:
                if (x_dev != x0_dev) SMM_CUDA(cudaMemcpyAsync(x_dev, x0_dev, sizeof(float) * (size_t)a->rows, cudaMemcpyDeviceToDevice, s));
                if (dist) { SMM_TRY(cg_init_dist(c, x0_dev)); iter = cg_iter_dist; }
                else { SMM_TRY(cg_init(c, x0_dev)); iter = cg_iter; }
                budget = max_it > 0 ? max_it : 0;                                                  // for-loop, H:2352
                break;
            case S_CG_IC0:
                if (!precond || smm_precond_kind(precond) != 1) { smm_set_error("ConjugateGradient(IC0): an IC0 preconditioner is required"); return SMM_E_INVALID; }
                if (x_dev != x0_dev) SMM_CUDA(cudaMemcpyAsync(x_dev, x0_dev, sizeof(float) * (size_t)a->rows, cudaMemcpyDeviceToDevice, s));
                SMM_TRY(pcg_init(c, x0_dev));
                iter = pcg_iter;
                budget = max_it > 0 ? max_it : 0;
                break;
            case S_BICGSYM:
                SMM_TRY(bicgsym_init(c)); iter = bicgsym_iter; budget = max_it > 1 ? max_it : 1;   // do-while, H:2047/2096
                break;
            case S_CGS:
                SMM_TRY(cgs_init(c)); iter = cgs_iter; budget = max_it > 1 ? max_it : 1;           // H:2131/2172
                break;
            case S_BICGSTAB:
                // the reference's BiCGStab is a template over anything with apply(rhs, x) (H:2191-2199): SGS from
                // getPreconditioner(), an IC0Preconditioner, or (extension) ILU(0) all go through the same sweeps
                SMM_TRY(stab_init(c)); iter = stab_iter; budget = max_it > 1 ? max_it : 1;         // H:2232/2277
                break;
        }
        launches += t_smm_launches - launches_before;
        int rc = SMM_OK;
        if (driver == SMM_DRIVER_PERSISTENT && !smm_cg_persistent_fits(a, c.x, c.r, c.p, c.ap)) driver = SMM_DRIVER_GRAPH_CHUNKED;
        if (budget > 0 && driver == SMM_DRIVER_PERSISTENT) {
            rc = smm_launch_cg_persistent(a, c.st, c.x, c.r, c.p, c.ap, s);
            launches += 1;
            if (rc == SMM_OK) rc = poll_state(c);
        } else if (budget > 0) {
            if (driver == SMM_DRIVER_GRAPH_WHILE) rc = run_graph_while(c, iter, &launches);
            else if (driver == SMM_DRIVER_GRAPH_CHUNKED) rc = run_graph_chunked(c, iter, budget, check_every, &launches);
            else rc = run_stream(c, iter, budget, check_every, &launches);
        } else {
            rc = poll_state(c);
        }
        if (rc != SMM_OK) return rc;
    } else {
        SMM_TRY(poll_state(c));
    }
    SMM_CUDA(cudaEventRecord(ws->ev1, s));
    SMM_CUDA(cudaEventSynchronize(ws->ev1));
    float ms = 0.f;
    SMM_CUDA(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));

    const SolveState* r = ws->state_host;
    int status = r->status;
    if (a->rows == 0) status = solver == S_CG ? SMM_SOLVER_SUCCESS : SMM_SOLVER_SUCCESS;
    if (hist_cap) {
        const int m = r->iterations < hist_cap ? r->iterations : hist_cap;
        if (m > 0) SMM_CUDA(cudaMemcpy(opts->history, ws->history, sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost));
    }
    if (info) {
        memset(info, 0, sizeof *info);
        info->status = status;
        info->iterations = r->iterations;
        info->residual = r->residual;
        info->precond_error = r->precond_error;
        info->seconds_solve = ms * 1e-3;
        info->seconds_total = ms * 1e-3;
        info->reduction_mode = c.mode;
        info->driver_mode = driver;
        info->kernel_launches = launches;
    }
    return SMM_OK;
}

// host-pointer front end: stage b, x0 in HBM, solve, copy x back
int solve_host(int solver, const smm_csr* a, const smm_precond* precond, const float* b, const float* x0, float* x,
               int max_iterations, float eps, const smm_solve_options* opts, smm_solve_info* info) {
    if (!a || (a->rows && (!b || !x || !x0))) { smm_set_error("solve: bad arguments"); return SMM_E_INVALID; }
    const auto t0 = std::chrono::steady_clock::now();
    SMM_CUDA(cudaSetDevice(a->device));
    smm_workspace* ws = nullptr;
    SMM_TRY(smm_workspace_get(a, &ws));
    SMM_TRY(smm_workspace_vectors(ws, 10, (size_t)a->rows));
    cudaStream_t s = smm_default_stream();
    const size_t bytes = sizeof(float) * (size_t)a->rows;
    float* d_b = ws->vec[0];
    float* d_x = ws->vec[1];
    float* d_x0 = d_x;
    if (a->rows) {
        SMM_CUDA(cudaMemcpyAsync(d_b, b, bytes, cudaMemcpyHostToDevice, s));
        SMM_CUDA(cudaMemcpyAsync(d_x, x0, bytes, cudaMemcpyHostToDevice, s));
    }
    smm_solve_info local;
    SMM_TRY(solve_dev(solver, a, precond, d_b, d_x0, d_x, max_iterations, eps, opts, &local, s));
    // ConjugateGradient returns before touching x when the initial residual already passes (H:2342-2344)
    const bool x_written = !((solver == S_CG || solver == S_CG_IC0) && local.iterations == 0) || x == x0;
    if (a->rows && x_written) {
        SMM_CUDA(cudaMemcpyAsync(x, d_x, bytes, cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
    }
    local.seconds_total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (info) *info = local;
    return SMM_OK;
}

}  // namespace

// every allocation a solve on `a` may need (workspace, work vectors, reference-order dot scratch, residual history), so that
// the solve itself allocates nothing: the single-process multi-GPU driver calls this on all devices before any kernel runs
int smm_solve_prepare(const smm_csr* a, const smm_solve_options* opts) {
    SMM_CUDA(cudaSetDevice(a->device));
    smm_workspace* ws = nullptr;
    SMM_TRY(smm_workspace_get(a, &ws));
    SMM_TRY(smm_workspace_vectors(ws, 10, (size_t)a->rows));
    if (opts && opts->reduction_mode != SMM_REDUCE_FAST) SMM_TRY(smm_dot_ref_prepare(a->rows, ws));
    const int hist_cap = opts && opts->history && opts->history_cap > 0 ? opts->history_cap : 0;
    if (hist_cap > ws->history_cap) {
        cudaFree(ws->history);
        ws->history = nullptr;
        SMM_CUDA(cudaMalloc(&ws->history, sizeof(float) * (size_t)hist_cap));
        ws->history_cap = hist_cap;
    }
    return SMM_OK;
}

int smm_solve_dist_cg_impl(smm_dist* d, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations, float eps,
                           const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s) {
    return solve_dev(S_CG, d->local, nullptr, b_dev, x0_dev, x_dev, maxIterations, eps, opts, info, s, d);
}
// solver: 1 BiCGSymmetric, 2 ConjugateGradientSquared, 3 BiCGStab (unpreconditioned); x is initial guess and result
int smm_solve_dist_impl(smm_dist* d, int solver, const float* b_dev, float* x_dev, int maxIterations, float eps,
                        const smm_solve_options* opts, smm_solve_info* info, cudaStream_t s) {
    const int kind = solver == 1 ? S_BICGSYM : (solver == 2 ? S_CGS : S_BICGSTAB);
    return solve_dev(kind, d->local, nullptr, b_dev, x_dev, x_dev, maxIterations, eps, opts, info, s, d);
}

extern "C" {

int smm_solve_cg(const smm_csr_t* a, const float* b, const float* x0, float* x, int maxIterations, float eps,
                 const smm_solve_options* opts, smm_solve_info* info) {
    return solve_host(S_CG, a, nullptr, b, x0, x, maxIterations, eps, opts, info);
}
int smm_solve_cg_ic0(const smm_csr_t* a, const smm_precond_t* ic0, const float* b, const float* x0, float* x, int maxIterations,
                     float eps, const smm_solve_options* opts, smm_solve_info* info) {
    return solve_host(S_CG_IC0, a, ic0, b, x0, x, maxIterations, eps, opts, info);
}
int smm_solve_cg_ic0_dev(const smm_csr_t* a, const smm_precond_t* ic0, const float* b_dev, const float* x0_dev, float* x_dev,
                         int maxIterations, float eps, const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return solve_dev(S_CG_IC0, a, ic0, b_dev, x0_dev, x_dev, maxIterations, eps, opts, info, stream ? (cudaStream_t)stream : smm_default_stream());
}
int smm_solve_bicgsym(const smm_csr_t* a, const float* b, float* x, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info) {
    return solve_host(S_BICGSYM, a, nullptr, b, x, x, maxIterations, eps, opts, info);
}
int smm_solve_cgs(const smm_csr_t* a, const float* b, float* x, int maxIterations, float eps,
                  const smm_solve_options* opts, smm_solve_info* info) {
    return solve_host(S_CGS, a, nullptr, b, x, x, maxIterations, eps, opts, info);
}
int smm_solve_bicgstab(const smm_csr_t* a, const smm_precond_t* precond, const float* b, float* x, int maxIterations,
                       float eps, const smm_solve_options* opts, smm_solve_info* info) {
    return solve_host(S_BICGSTAB, a, precond, b, x, x, maxIterations, eps, opts, info);
}

static inline cudaStream_t pick_stream(void* stream) { return stream ? (cudaStream_t)stream : smm_default_stream(); }

int smm_solve_cg_dev(const smm_csr_t* a, const float* b_dev, const float* x0_dev, float* x_dev, int maxIterations,
                     float eps, const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return solve_dev(S_CG, a, nullptr, b_dev, x0_dev, x_dev, maxIterations, eps, opts, info, pick_stream(stream));
}
int smm_solve_bicgsym_dev(const smm_csr_t* a, const float* b_dev, float* x_dev, int maxIterations, float eps,
                          const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return solve_dev(S_BICGSYM, a, nullptr, b_dev, x_dev, x_dev, maxIterations, eps, opts, info, pick_stream(stream));
}
int smm_solve_cgs_dev(const smm_csr_t* a, const float* b_dev, float* x_dev, int maxIterations, float eps,
                      const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return solve_dev(S_CGS, a, nullptr, b_dev, x_dev, x_dev, maxIterations, eps, opts, info, pick_stream(stream));
}
int smm_solve_bicgstab_dev(const smm_csr_t* a, const smm_precond_t* precond, const float* b_dev, float* x_dev,
                           int maxIterations, float eps, const smm_solve_options* opts, smm_solve_info* info, void* stream) {
    return solve_dev(S_BICGSTAB, a, precond, b_dev, x_dev, x_dev, maxIterations, eps, opts, info, pick_stream(stream));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// measurement hook: per-kernel device times of one fused CG iteration (bench.py's roofline line).
// Launches the SAME three kernels the solver's iteration graph holds, `reps` times each, on `stream`, bracketed
// by CUDA events.  Vectors are the handle's own work vectors (contents irrelevant for timing, kept finite).
// ---------------------------------------------------------------------------------------------------
// dist != null: the kernels of the multi-GPU iteration on this rank's rows (p inside the extended vector, halo stores fused into
// the x,p update, interior rows first in the SpMV).  Only meaningful with SMM_B200_DIST_DEBUG=7 set when the handle was created:
// the timed launches of one kernel follow each other without the peers taking part, so nothing may wait for them.
int smm_profile_cg_iteration_impl(const smm_csr_t* a, smm_dist* dist, int reps, float* ms_spmv, float* ms_xr, float* ms_p, void* stream) {
    if (!a || reps < 1) return SMM_E_INVALID;
    SMM_CUDA(cudaSetDevice(a->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    smm_workspace* ws = nullptr;
    SMM_TRY(smm_workspace_get(a, &ws));
    SMM_TRY(smm_workspace_vectors(ws, 7, (size_t)a->rows));
    Ctx c;
    c.a = a; c.ws = ws; c.s = s; c.st = ws->state; c.n = a->rows;
    c.dist = dist;
    float** w = ws->vec + 3;
    c.x = ws->vec[1]; c.r = w[0]; c.p = w[1]; c.ap = w[2];
    if (dist) c.p = dist->ext + dist->own_off;
    const size_t bytes = sizeof(float) * (size_t)a->rows;
    SMM_CUDA(cudaMemsetAsync(c.x, 0, bytes, s));
    SMM_CUDA(cudaMemsetAsync(c.r, 0, bytes, s));
    SMM_CUDA(cudaMemsetAsync(c.p, 0, bytes, s));
    SMM_CUDA(cudaMemsetAsync(c.st, 0, sizeof(SolveState), s));
    float* outs[3] = {ms_spmv, ms_xr, ms_p};
    for (int k = 0; k < 3; ++k) {
        for (int rep = -2; rep < reps; ++rep) {                // two untimed warm-up launches
            if (rep == 0) SMM_CUDA(cudaEventRecord(ws->ev0, s));
            if (k == 0) SMM_TRY(spmv(c, SMM_OP_ASSIGN, nullptr, dist ? dist->ext : c.p, c.ap, RED_OUT_AUX, FIN_CG_ALPHA, c.p));
            else if (k == 1) SMM_TRY(vec(c, VEC_CG_R, FIN_STORE, {c.r, c.ap}, {c.r}));
            else { SMM_TRY(vec(c, VEC_CG_PX, FIN_NONE, {c.p, c.r, c.x}, {c.p, c.x}, dist != nullptr)); if (dist) SMM_TRY(dist_after_p(c)); }
        }
        SMM_CUDA(cudaEventRecord(ws->ev1, s));
        SMM_CUDA(cudaEventSynchronize(ws->ev1));
        float ms = 0.f;
        SMM_CUDA(cudaEventElapsedTime(&ms, ws->ev0, ws->ev1));
        if (outs[k]) *outs[k] = ms / reps;
    }
    return SMM_OK;
}

extern "C" int smm_profile_cg_iteration(const smm_csr_t* a, int reps, float* ms_spmv, float* ms_xr, float* ms_p, void* stream) {
    return smm_profile_cg_iteration_impl(a, nullptr, reps, ms_spmv, ms_xr, ms_p, stream);
}
