// epilogue.cuh -- the scalar part of every Krylov step, executed on the device by ONE thread (thread 0 of the
// CTA that finishes a grid-wide reduction).  This is where the reference's host-side control flow lives on the
// GPU: alpha/beta/omega, the stopping tests, the DIVERGED heuristics and the iteration counter.  Comparisons are
// written exactly as in the reference so that NaN/Inf take the same branches.
//
// All float ops use the round-to-nearest intrinsics so nvcc cannot contract or reassociate them.
#pragma once
#include "dist_device.cuh"
#include "smm_internal.cuh"

enum ReduceShape {
    RED_NONE = 0,
    RED_OUT_OUT = 1,            // t0 = out.out
    RED_OUT_AUX = 2,            // t0 = out.aux
    RED_OUT_AUX_OUT_OUT = 3,    // t0 = out.aux, t1 = out.out
};

enum FinishKind {
    FIN_NONE = 0,
    FIN_STORE,                  // scratch[0] = t0, scratch[1] = t1 (plain dot products)
    FIN_CG_INIT,                // H:2341-2347
    FIN_RR_INIT,                // H:2043 / H:2128 / H:2231
    FIN_CG_ALPHA,               // H:2354-2358
    FIN_CG_UPDATE,              // H:2369-2382 (+ loop bound H:2352, H:2397)
    FIN_BICGSYM_ALPHA,          // H:2049-2059
    FIN_BICGSYM_UPDATE,         // H:2075-2096
    FIN_ALPHA_R0,               // H:2133-2135 / H:2243-2244
    FIN_CGS_UPDATE,             // H:2152-2154, H:2169-2172
    FIN_BICGSTAB_OMEGA,         // H:2259-2261
    FIN_BICGSTAB_UPDATE,        // H:2268-2277  (t0 = sum r^2, t1 = r.r0)
    FIN_STASH0,                 // scratch[0] = t0 (first half of a two-kernel reduction)
    FIN_BICGSTAB_UPDATE_STASHED,// as FIN_BICGSTAB_UPDATE with sum r^2 = scratch[0], r.r0 = t0
    FIN_PCG_INIT,               // H:2442-2454  (t0 = r.z, t1 = r.r)
    FIN_PCG_ALPHA,              // H:2462-2466
    FIN_PCG_UPDATE,             // H:2483-2501  (t0 = r.z, t1 = r.r)
};

#ifdef __CUDACC__
__device__ __forceinline__ void smm_push_history(SolveState* st, float v) {
    if (st->history != nullptr && st->iterations < st->history_cap) st->history[st->iterations] = v;
}

// `if (iterations > maxIterations) return MAX_ITERATIONS_REACHED; return SUCCESS;` after the do-while loops: only
// reachable with maxIterations <= 0 (the body still runs once), but then the reference does return it.
__device__ __forceinline__ int smm_do_while_status(const SolveState* st) {
    return st->iterations > st->max_iterations ? SMM_SOLVER_MAX_ITERATIONS_REACHED : SMM_SOLVER_SUCCESS;
}

__device__ __forceinline__ void smm_finish(int kind, SolveState* st, float t0, float t1) {
    if (kind == FIN_NONE) return;
    if (st->comm != nullptr) {
        dist_allreduce2(st->comm, t0, t1);                     // every rank now holds the same totals
        if (st->comm->error) { st->done = 1; st->precond_error |= 8; return; }
    }
    switch (kind) {
        case FIN_STORE:
            st->scratch[0] = t0;
            st->scratch[1] = t1;
            break;
        case FIN_CG_INIT:
            st->rr = t0;                                       // residualNormSquared = r * r
            st->residual = t0;
            if (st->eps2 > t0) { st->done = 1; st->status = SMM_SOLVER_SUCCESS; }            // H:2342-2344
            else if (st->iterations >= st->max_iterations) { st->done = 1; st->status = SMM_SOLVER_MAX_ITERATIONS_REACHED; }
            break;
        case FIN_RR_INIT:
            st->rr = t0;
            st->residual = t0;
            break;
        case FIN_CG_ALPHA:
            st->denom = t0;                                    // pAp = Ap * p
            st->alpha = __fdiv_rn(st->rr, t0);                 // H:2358 (pAp == 0 only asserted: inf/NaN propagate)
            break;
        case FIN_CG_UPDATE: {
            const float new_rr = t0;                           // H:2369
            smm_push_history(st, new_rr);
            st->iterations += 1;
            st->residual = new_rr;
            // x_owed: with the two-pass iteration (VEC_CG_R / VEC_CG_PX) the x update of this iteration comes after this test;
            // the flag lets that one kernel run although the solve is over (harmless for VEC_CG_XR, which has updated x already)
            if (st->eps2 > new_rr) { st->done = 1; st->x_owed = 1; st->status = SMM_SOLVER_SUCCESS; break; }  // H:2377-2379
            st->beta = __fdiv_rn(new_rr, st->rr);              // H:2381
            st->rr = new_rr;
            if (st->iterations >= st->max_iterations) { st->done = 1; st->x_owed = 1; st->status = SMM_SOLVER_MAX_ITERATIONS_REACHED; }  // H:2397
            break;
        }
        case FIN_BICGSYM_ALPHA:
            st->denom = t0;
            if (st->eps > fabsf(t0) && st->rr > 1.0f) { st->done = 1; st->status = SMM_SOLVER_DIVERGED; break; }  // H:2056-2058
            st->alpha = __fdiv_rn(st->rr, t0);                 // H:2059
            break;
        case FIN_BICGSYM_UPDATE: {
            const float new_rr = t0;                           // H:2075
            if (new_rr > 1.0f && st->rr < st->eps) {           // H:2079-2081
                st->done = 1; st->x_owed = 1; st->status = SMM_SOLVER_DIVERGED; st->residual = new_rr; break;   // x has been updated by then (H:2061-2067)
            }
            st->beta = __fdiv_rn(new_rr, st->rr);              // H:2082
            st->rr = new_rr;                                   // H:2094
            smm_push_history(st, new_rr);
            st->iterations += 1;                               // H:2095
            st->residual = new_rr;
            if (!(new_rr > st->eps2 && st->iterations < st->max_iterations)) {                 // H:2096
                st->done = 1; st->x_owed = 1; st->status = smm_do_while_status(st);            // H:2098-2101
            }
            break;
        }
        case FIN_ALPHA_R0:
            st->denom = t0;                                    // ap * r0
            st->alpha = __fdiv_rn(st->rr, t0);                 // H:2135 / H:2244
            break;
        case FIN_CGS_UPDATE: {
            const float new_rr0 = t0;                          // r * r0, H:2152
            st->beta = __fdiv_rn(new_rr0, st->rr);             // H:2154
            st->rr = new_rr0;                                  // H:2169
            smm_push_history(st, t1);
            st->iterations += 1;                               // H:2170
            st->res2 = t1;                                     // r * r, H:2171
            st->residual = t1;
            if (!(t1 > st->eps2 && st->iterations < st->max_iterations)) {                     // H:2172
                st->done = 1; st->status = smm_do_while_status(st);                            // H:2174-2177
            }
            break;
        }
        case FIN_BICGSTAB_OMEGA:
            st->as_s = t0;
            st->as_as = t1;
            st->omega = __fdiv_rn(t0, t1);                     // H:2259-2261
            break;
        case FIN_STASH0:
            st->scratch[0] = t0;
            break;
        case FIN_BICGSTAB_UPDATE_STASHED:
            t1 = t0;
            t0 = st->scratch[0];
            // fall through
        case FIN_BICGSTAB_UPDATE: {
            const float res = __fsqrt_rn(t0);                  // H:2268
            const float new_rr0 = t1;                          // H:2269
            st->beta = __fdiv_rn(__fmul_rn(new_rr0, st->alpha), __fmul_rn(st->rr, st->omega));  // H:2271
            st->rr = new_rr0;                                  // H:2275
            smm_push_history(st, res);
            st->iterations += 1;                               // H:2276
            st->res2 = t0;
            st->residual = res;
            if (!(res > st->eps && st->iterations < st->max_iterations)) {                     // H:2277
                st->done = 1; st->status = smm_do_while_status(st);                            // H:2279-2282
            }
            break;
        }
        case FIN_PCG_INIT:
            st->rr = t0;                                       // rz
            st->res2 = t1;                                     // residualNormSquared
            st->residual = t1;
            if (st->eps2 > t1) { st->done = 1; st->status = SMM_SOLVER_SUCCESS; }             // H:2449-2451
            else if (st->iterations >= st->max_iterations) { st->done = 1; st->status = SMM_SOLVER_MAX_ITERATIONS_REACHED; }
            break;
        case FIN_PCG_ALPHA:
            st->denom = t0;                                    // pAp
            st->alpha = __fdiv_rn(st->rr, t0);                 // alpha = rz / pAp, H:2466
            break;
        case FIN_PCG_UPDATE: {
            smm_push_history(st, t1);
            st->iterations += 1;
            st->res2 = t1;                                     // r * r, H:2484
            st->residual = t1;
            if (st->eps2 > t1) { st->done = 1; st->status = SMM_SOLVER_SUCCESS; break; }      // H:2485-2487
            st->beta = __fdiv_rn(t0, st->rr);                  // newRZ / rz, H:2488
            st->rr = t0;                                       // H:2501
            if (st->iterations >= st->max_iterations) { st->done = 1; st->status = SMM_SOLVER_MAX_ITERATIONS_REACHED; }  // H:2504
            break;
        }
        default: break;
    }
}
#endif
