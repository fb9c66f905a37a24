// spmv.cu -- CSR SpMV for sm_100a: out[row] = op(lhs[row], sum_k values[k] * mult[positions[k]])
//
// Replaces CSRMatrix<float>::rMultOp / rMult / rMultAdd / rMultSub of the reference (H:1458-1515), and fuses the
// dot products the Krylov solvers take right after it (H:2354, H:2133, H:2243, H:2259-2261, H:2152+H:2171,
// H:2341) into its epilogue, so that an iteration never re-reads the SpMV result for a reduction.
//
// Work decomposition ("nnz-chunked row ranges"): CTA q owns the rows whose FIRST entry lies in
// [q*CHUNK, (q+1)*CHUNK) (block_row[], found once per matrix by binary search).  Every CTA therefore streams
// about CHUNK entries whatever the row-length distribution, and chooses per range, from its row-length
// statistics (rows in range, entries in range):
//   * stream path (many short rows; stencils, the bulk of a power-law matrix): all threads load values/positions
//     with 128-bit coalesced streaming loads, gather mult[] through L1/L2, and stage the products in shared
//     memory; then one thread per row adds its products LEFT TO RIGHT -- the reference's accumulation order, so
//     these rows are bit-identical to the reference.  Rows longer than SPMV_LONG_IN_STREAM inside such a range are
//     summed by a warp (shuffle reduction) instead.
//   * row path (few long rows): warp per row with coalesced strided loads and a shuffle reduction; rows longer
//     than SPMV_CTA_ROW are reduced by the whole CTA.
//   * exact mode: every row is accumulated left to right by one thread (parity runs).
//
// Matrices whose rows are all short and of similar length (stencils: BASELINE configs 1, 2, 3, 5) take a second,
// Blackwell-specific kernel instead (spmv_rows_kernel): a persistent CTA per SM slot streams fixed groups of rows;
// the TMA engine (cp.async.bulk + mbarrier, 3 stages) copies each group's values/positions window into shared
// memory while the previous groups are being multiplied, and V lanes per row (V = 1, 2, 4, 8 chosen from the mean
// row length; V = 1 for stencils) walk their row out of shared memory, so that consecutive lanes gather CONSECUTIVE
// entries of mult[] (one or two 128-byte lines per warp-wide gather instead of one line per distinct stencil
// offset), and accumulate in registers -- for V = 1 in the reference's left-to-right order, bit-identical.
// Algorithmic bytes per launch: 8*nnz (values+positions) + 4*(rows+1) (start) + 4*cols (mult, gathered once)
// + 4*rows (out) [+ 4*rows lhs for ADD/SUB] [+ 4*rows per fused-dot operand].
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "smm_internal.cuh"

struct HaloWaitDev;   // dist_device.cuh

namespace {

struct SpmvParams {
    const int32_t* __restrict__ start;
    const int32_t* __restrict__ positions;
    const float* __restrict__ values;
    const int32_t* __restrict__ block_row;
    const int32_t* __restrict__ row_perm;   // rows of every chunk, longest first (null: in index order)
    const uint16_t* __restrict__ start16;   // rows kernel: [groups][R + 1] row starts relative to the group's window start (a0)
    int rows;
    int nnz;
    int nnz_alloc;        // entries readable in positions/values (TMA windows are rounded up to 16 bytes)
    int op;
    int exact;
    const float* lhs;     // may alias out
    const float* mult;    // never aliases out
    float* out;
    float* copy1;
    float* copy2;
    float* copy3;
    const float* aux;
    int reduce;           // ReduceShape
    int finish;           // FinishKind
    SolveState* state;
    float* partials;
    size_t partials_stride;
    unsigned int* ticket;
    const HaloWaitDev* halo;         // multi-GPU: the rows that read halo entries wait for the peers' flags (rows kernel only)
};

}  // namespace

#include "epilogue.cuh"
#include "vec_functors.cuh"

namespace {

struct RowWriter {
    const SpmvParams& P;
    float acc0 = 0.0f, acc1 = 0.0f;
    __device__ __forceinline__ explicit RowWriter(const SpmvParams& p) : P(p) {}
    __device__ __forceinline__ void operator()(int row, float dot) {
        float o;
        if (P.op == SMM_OP_ASSIGN) {
            o = dot;                                          // vectorMultFunctor, H:1284-1286
        } else {
            const float l = P.lhs[row];
            o = (P.op == SMM_OP_ADD) ? __fadd_rn(l, dot) : __fsub_rn(l, dot);   // H:1509, H:1514
        }
        P.out[row] = o;
        if (P.copy1) P.copy1[row] = o;
        if (P.copy2) P.copy2[row] = o;
        if (P.copy3) P.copy3[row] = o;
        switch (P.reduce) {
            case RED_OUT_OUT: acc0 = fmaf(o, o, acc0); break;
            case RED_OUT_AUX: acc0 = fmaf(o, P.aux[row], acc0); break;
            case RED_OUT_AUX_OUT_OUT: acc0 = fmaf(o, P.aux[row], acc0); acc1 = fmaf(o, o, acc1); break;
            default: break;
        }
    }
};

// out[row] = dot with no lhs and no extra copies: the SpMV of every Krylov iteration (rMult, H:1501-1505)
struct PlainWriter {
    const SpmvParams& P;
    float acc0 = 0.0f, acc1 = 0.0f;
    __device__ __forceinline__ explicit PlainWriter(const SpmvParams& p) : P(p) {}
    __device__ __forceinline__ void operator()(int row, float o) {
        P.out[row] = o;
        if (P.reduce == RED_OUT_AUX) acc0 = fmaf(o, P.aux[row], acc0);
        else if (P.reduce == RED_OUT_AUX_OUT_OUT) { acc0 = fmaf(o, P.aux[row], acc0); acc1 = fmaf(o, o, acc1); }
        else if (P.reduce == RED_OUT_OUT) acc0 = fmaf(o, o, acc0);
    }
};

// Rows [r0, r1) straight from global memory: exact mode = one thread per row, left to right; otherwise warp per row
// (coalesced strided loads + shuffle reduction) and the whole CTA for rows longer than SPMV_CTA_ROW.
// Must be called by every thread of the CTA (it contains block-wide barriers).
template <class Writer>
__device__ __forceinline__ void row_path(const SpmvParams& P, int r0, int r1, Writer& write, float* red_sh) {
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;
    if (P.exact) {
        for (int r = r0 + tid; r < r1; r += nthreads) {
            const int s = P.start[r], e = P.start[r + 1];
            float dot = 0.0f;
            for (int k = s; k < e; ++k) dot = __fadd_rn(__fmul_rn(P.values[k], __ldg(P.mult + P.positions[k])), dot);
            write(r, dot);
        }
        return;
    }
    bool any_cta_row = false;
    for (int r = r0 + warp; r < r1; r += nwarps) {
        const int s = P.start[r], e = P.start[r + 1];
        if (e - s > SPMV_CTA_ROW) { any_cta_row = true; continue; }
        float acc = 0.0f;
        int k = s + lane;
        for (; k + 32 < e; k += 64) {                         // two independent gathers in flight
            const int c0 = ldg_stream_i(P.positions + k), c1 = ldg_stream_i(P.positions + k + 32);
            const float a0v = ldg_stream_f(P.values + k), a1v = ldg_stream_f(P.values + k + 32);
            acc = fmaf(a0v, __ldg(P.mult + c0), acc);
            acc = fmaf(a1v, __ldg(P.mult + c1), acc);
        }
        if (k < e) acc = fmaf(ldg_stream_f(P.values + k), __ldg(P.mult + ldg_stream_i(P.positions + k)), acc);
        acc = warp_sum(acc);
        if (lane == 0) write(r, acc);
    }
    if (__syncthreads_or(any_cta_row)) {                      // uniform: rescan for CTA-wide rows
        for (int r = r0; r < r1; ++r) {
            const int s = P.start[r], e = P.start[r + 1];
            if (e - s <= SPMV_CTA_ROW) continue;
            float acc = 0.0f;
            const int sa = (s + 3) & ~3;                      // aligned body, scalar head/tail
            const int ea = e & ~3;
            if (tid < sa - s) acc = fmaf(P.values[s + tid], __ldg(P.mult + P.positions[s + tid]), acc);
            if (tid < e - ea) acc = fmaf(P.values[ea + tid], __ldg(P.mult + P.positions[ea + tid]), acc);
            const int4* pos4 = reinterpret_cast<const int4*>(P.positions + sa);
            const float4* val4 = reinterpret_cast<const float4*>(P.values + sa);
            const int nvec = (ea - sa) >> 2;
            for (int v = tid; v < nvec; v += nthreads) {
                const int4 c = ldg_stream_i4(pos4 + v);
                const float4 a = ldg_stream_f4(val4 + v);
                acc = fmaf(a.x, __ldg(P.mult + c.x), acc);
                acc = fmaf(a.y, __ldg(P.mult + c.y), acc);
                acc = fmaf(a.z, __ldg(P.mult + c.z), acc);
                acc = fmaf(a.w, __ldg(P.mult + c.w), acc);
            }
            float v1[1] = {acc};
            __syncthreads();
            block_sum<1>(v1, red_sh);
            if (tid == 0) write(r, v1[0]);
            __syncthreads();
        }
    }
}

// DEPTH: 128-bit index/value vectors a thread has in flight before its gathers (4 * DEPTH gathers of mult[] per thread);
// HINTS: L2 eviction priorities -- the streamed values/positions are marked evict-first and the gathered vector evict-last, so
// that on matrices whose gathers have no locality (power-law rows, config 4) the 8 * nnz bytes passing through do not push
// the 4 * cols bytes that are gathered again and again out of L2.
template <int DEPTH, bool HINTS>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_kernel(const SpmvParams P) {
    smm_pdl_wait();
    if (P.state != nullptr && P.state->done) {
        // an SpMV after the end of the solve: the x update a two-pass CG iteration still owed (VEC_CG_PX) has run by now
        if (blockIdx.x == 0 && threadIdx.x == 0 && P.state->x_owed) P.state->x_owed = 0;
        return;
    }

    __shared__ __align__(16) float prod[SPMV_CAP];
    __shared__ float red_sh[96];
    __shared__ int sh_flag;
    __shared__ int sh_long;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = SPMV_THREADS / 32;
    const int r0 = P.block_row[blockIdx.x];
    const int r1 = P.block_row[blockIdx.x + 1];
    RowWriter write(P);

    if (r1 > r0) {
        const int k0 = P.start[r0];
        const int k1 = P.start[r1];
        const int a0 = k0 & ~3;                               // 16-byte aligned start of the staged window
        const int span = k1 - a0;
        const int nrows = r1 - r0;
        const bool fits = span <= SPMV_CAP;
        if (fits && (P.exact || nrows >= SPMV_MIN_STREAM_ROWS)) {
            // ---------------- stream path ----------------
            if (tid == 0) sh_long = 0;
            const int4* pos4 = reinterpret_cast<const int4*>(P.positions + a0);
            const float4* val4 = reinterpret_cast<const float4*>(P.values + a0);
            const int nvec = span >> 2;                       // full vectors inside [a0, k1)
            // DEPTH vectors per thread in flight before the dependent gathers
            unsigned long long pol_stream = 0ull, pol_keep = 0ull;
            if (HINTS) {
                asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
                asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            }
            auto load_c = [&](const int4* p) {
                if (!HINTS) return ldg_stream_i4(p);
                int4 r;
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol_stream));
                return r;
            };
            auto load_a = [&](const float4* p) {
                if (!HINTS) return ldg_stream_f4(p);
                float4 r;
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol_stream));
                return r;
            };
            auto load_x = [&](const int c) {
                if (!HINTS) return __ldg(P.mult + c);
                float r;
                asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(P.mult + c), "l"(pol_keep));
                return r;
            };
            for (int v = tid; v < nvec; v += DEPTH * SPMV_THREADS) {
                int4 c[DEPTH];
                float4 a[DEPTH];
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const int vd = v + d * SPMV_THREADS;
                    if (vd < nvec) { c[d] = load_c(pos4 + vd); a[d] = load_a(val4 + vd); }
                    else { c[d] = make_int4(0, 0, 0, 0); a[d] = make_float4(0.f, 0.f, 0.f, 0.f); }
                }
                float4 x[DEPTH];
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const bool in = v + d * SPMV_THREADS < nvec;
                    x[d] = in ? make_float4(load_x(c[d].x), load_x(c[d].y), load_x(c[d].z), load_x(c[d].w)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int d = 0; d < DEPTH; ++d) {
                    const int vd = v + d * SPMV_THREADS;
                    if (vd < nvec) reinterpret_cast<float4*>(prod)[vd] = make_float4(__fmul_rn(a[d].x, x[d].x), __fmul_rn(a[d].y, x[d].y), __fmul_rn(a[d].z, x[d].z), __fmul_rn(a[d].w, x[d].w));
                }
            }
            {   // tail of the window (fewer than 4 entries)
                const int k = a0 + (nvec << 2) + tid;
                if (k < k1) prod[k - a0] = __fmul_rn(ldg_stream_f(P.values + k), __ldg(P.mult + ldg_stream_i(P.positions + k)));
            }
            __syncthreads();
            // One thread per row, left to right.  With row_perm the rows of the range come longest first: the 32 rows of a
            // warp have similar lengths (a warp costs its longest row), and the rows too long for one thread lead the list.
            const bool sorted = P.row_perm != nullptr && nrows <= SPMV_PERM_MAX_ROWS;
            bool saw_long = false;
            for (int i = tid; i < nrows; i += SPMV_THREADS) {
                const int r = sorted ? P.row_perm[r0 + i] : r0 + i;
                const int s = P.start[r] - a0, e = P.start[r + 1] - a0;
                if (!P.exact && e - s > SPMV_LONG_IN_STREAM) { saw_long = true; continue; }
                float dot = 0.0f;                             // H:1484 ; empty row -> op(lhs, 0), H:1479-1483
                for (int j = s; j < e; ++j) dot = __fadd_rn(prod[j], dot);   // H:1485-1489, `val*x + dot`, two roundings
                write(r, dot);
            }
            if (saw_long) sh_long = 1;
            __syncthreads();
            if (sh_long) {
                for (int i = warp; i < nrows; i += NWARPS) {
                    const int r = sorted ? P.row_perm[r0 + i] : r0 + i;
                    const int s = P.start[r] - a0, e = P.start[r + 1] - a0;
                    if (e - s <= SPMV_LONG_IN_STREAM) { if (sorted) break; continue; }
                    float acc = 0.0f;
                    for (int j = s + lane; j < e; j += 32) acc += prod[j];
                    acc = warp_sum(acc);
                    if (lane == 0) write(r, acc);
                }
            }
        } else {
            row_path(P, r0, r1, write, red_sh);
        }
    }

    smm_pdl_trigger();
    if (P.reduce != RED_NONE) {
        float v[2] = {write.acc0, write.acc1};
        __syncthreads();
        if (grid_sum_last_block<2>(v, P.partials, P.partials_stride, P.ticket, red_sh, &sh_flag)) {
            if (tid == 0) smm_finish(P.finish, P.state, v[0], v[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// spmv_rows_kernel: persistent, TMA-staged, V lanes per row (see the file header)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is signalled on the mbarrier in bytes
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int ROWS_MAX_STAGES = 8;
constexpr int ROWS_CONSUMERS = SPMV_THREADS;                      // 8 consumer warps
constexpr int ROWS_THREADS = SPMV_THREADS + 32;                   // + 1 producer warp

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One row (or a V-lane slice of it) out of a values/positions window.  Up to eight entries are fetched and their
// mult[] gathers issued together (one latency round for a stencil row), then added in the order of the entries
// (V = 1: the reference's left-to-right sum with two roundings per term, H:1484-1489).
// exactly L entries, no predicates (the interior rows of a stencil)
// COH: the operand vector is being written by peer GPUs while this kernel runs (halo rows of a multi-GPU SpMV, after the
// flags have been acquired): gather past L1 (ld.global.cg), whose lines may predate the peers' stores
template <bool COH>
__device__ __forceinline__ float gather(const float* __restrict__ mult, const int c) { return COH ? __ldcg(mult + c) : __ldg(mult + c); }

template <int L, bool COH, class VP, class CP>
__device__ __forceinline__ float row_dot_fixed(const VP vs, const CP cs, const float* __restrict__ mult, const int j) {
    int c[L];
    float v[L], x[L];
#pragma unroll
    for (int k = 0; k < L; ++k) { c[k] = cs[j + k]; v[k] = vs[j + k]; }
#pragma unroll
    for (int k = 0; k < L; ++k) x[k] = gather<COH>(mult, c[k]);
    float dot = 0.0f;
#pragma unroll
    for (int k = 0; k < L; ++k) dot = __fadd_rn(__fmul_rn(v[k], x[k]), dot);
    return dot;
}

template <int V, bool COH, class VP, class CP>
__device__ __forceinline__ float row_dot(const VP vs, const CP cs, const float* __restrict__ mult, int j, const int e) {
    if (V == 1) {                                                 // warp-uniform row length: same sum, fewer instructions
        const int len = e - j;
        if (__all_sync(0xffffffffu, len == 7)) return row_dot_fixed<7, COH>(vs, cs, mult, j);
        if (__all_sync(0xffffffffu, len == 5)) return row_dot_fixed<5, COH>(vs, cs, mult, j);
    }
    float dot = 0.0f;
    while (j < e) {
        int c[8];
        float v[8], x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool in = j + k * V < e;
            c[k] = in ? cs[j + k * V] : -1;
            v[k] = in ? vs[j + k * V] : 0.0f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = c[k] >= 0 ? gather<COH>(mult, c[k]) : 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (c[k] >= 0) dot = __fadd_rn(__fmul_rn(v[k], x[k]), dot);
        }
        j += 8 * V;
    }
    return dot;
}

// Warp-specialised persistent kernel.  Warp 8 is the producer: for each of this CTA's row groups it waits until
// the slot is free (empty barrier), then has the TMA engine copy the group's values/positions window into the
// slot, completion counted in bytes on the slot's full barrier.  Warps 0..7 are consumers: wait for the slot, take
// one row per V lanes out of shared memory, release the slot.  No CTA-wide barrier inside the loop.
// One pass of a (virtual) CTA over its row groups first, first + G, ... : the body of spmv_rows_kernel, also run -- once per
// virtual CTA and iteration -- by the persistent CG kernel below.  ring_s / ring_k: the thread's position in the shared-memory
// ring (slot; uses of the slot for the producer, phase parity for a consumer), carried from one pass to the next.
template <int V, bool HALO, class Writer>
__device__ __forceinline__ void rows_sweep(const SpmvParams& P, const int cap, const int nchunks, const int stages, float* vals_s, int* cols_s,
                                           uint64_t* full, uint64_t* empty, int* win, Writer& write, const int first, const int G, int& ring_s, uint32_t& ring_k) {
    constexpr int R = ROWS_CONSUMERS / V;                         // rows per group
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int my_chunks = first < nchunks ? (nchunks - first + G - 1) / G : 0;
    // Multi-GPU: groups [0, c_lo) and [c_hi, nchunks) hold rows that read halo entries.  The groups are walked in the order
    // c_lo .. nchunks-1, 0 .. c_lo-1 (group = position + c_lo, wrapped), i.e. the interior first, so that the peers' pushes
    // travel while the bulk of the rows is being multiplied; a warp spins on the flags only when it reaches a boundary group.
    int c_lo = 0, c_hi = nchunks;
    if (HALO) {
        c_lo = min(nchunks, (P.halo->row_lo + R - 1) / R);
        c_hi = max(c_lo, P.halo->row_hi / R);
    }
    auto group_of = [&](const int j) {
        if (!HALO) return j;
        const int q = j + c_lo;
        return q >= nchunks ? q - nchunks : q;
    };

    if (warp == ROWS_CONSUMERS / 32) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int k0 = 0, k1 = 0;
            if (my_chunks > 0) {
                const int rb = group_of(first) * R, re = min(rb + R, P.rows);
                k0 = P.start[rb]; k1 = P.start[re];
            }
            int s = ring_s;                                        // ring slot and how often it has been used before
            uint32_t use = ring_k;
            for (int it = 0; it < my_chunks; ++it) {
                int n0 = 0, n1 = 0;                                // next group's window, fetched ahead of the wait
                if (it + 1 < my_chunks) {
                    const int rb = group_of(first + (it + 1) * G) * R, re = min(rb + R, P.rows);
                    n0 = P.start[rb]; n1 = P.start[re];
                }
                if (use > 0) mbar_wait(&empty[s], (use - 1) & 1u);
                const int a0 = k0 & ~3;
                const int span = k1 - a0;
                // the copy is rounded up to 16 bytes: it must stay inside the allocation (adopted arrays may not be padded)
                const bool staged = span > 0 && span <= cap && a0 + ((span + 3) & ~3) <= P.nnz_alloc;
                win[2 * s] = a0;
                win[2 * s + 1] = staged ? 1 : 0;
                if (staged) {
                    const uint32_t bytes = (uint32_t)(((span + 3) & ~3) * 4);
                    mbar_expect_tx(&full[s], 2 * bytes);
                    tma_load_1d(vals_s + (size_t)s * cap, P.values + a0, bytes, &full[s]);
                    tma_load_1d(cols_s + (size_t)s * cap, P.positions + a0, bytes, &full[s]);
                } else {
                    mbar_arrive(&full[s]);                         // nothing to copy: consumers read global memory
                }
                k0 = n0; k1 = n1;
                if (++s == stages) { s = 0; ++use; }
            }
            ring_s = s; ring_k = use;
        }
    } else {
        // ---------------- consumers ----------------
        const int sub = tid % V;                                  // lane inside the row's lane group
        const int rloc = tid / V;                                 // row inside the group
        int my_s = 0, my_e = 0;
        if (my_chunks > 0) {
            const uint16_t* s16 = P.start16 + (size_t)group_of(first) * (R + 1) + rloc;    // relative to the group's window start
            my_s = s16[0]; my_e = s16[1];
        }
        int s = ring_s;
        uint32_t phase = ring_k;
        int j = first;
        bool halo_here = !HALO;                                   // the peers' halo entries have been acquired by this warp
        for (int it = 0; it < my_chunks; ++it, j += G) {
            const int q = group_of(j);
            int nx_s = 0, nx_e = 0;                               // next group's row bounds, fetched ahead
            if (it + 1 < my_chunks) {
                const uint16_t* s16 = P.start16 + (size_t)group_of(j + G) * (R + 1) + rloc;
                nx_s = s16[0]; nx_e = s16[1];
            }
            const int row = q * R + rloc;
            const bool boundary = HALO && (q < c_lo || q >= c_hi);
            if (boundary && !halo_here) {
                if (lane == 0 && !dist_halo_wait(P.halo) && P.state != nullptr) { P.state->done = 1; P.state->precond_error |= 8; }
                __syncwarp();
                halo_here = true;
            }
            mbar_wait(&full[s], phase);
            const int a0 = win[2 * s];
            float dot;
            // my_s / my_e are relative to a0, the start of the group's window (start16: 2 bytes per row instead of the 4 of start[]);
            // rows past the end have my_s == my_e: their lanes fall through and only join the shuffles
            if (HALO && boundary) {
                if (win[2 * s + 1]) dot = row_dot<V, true>(vals_s + (size_t)s * cap, cols_s + (size_t)s * cap, P.mult, my_s + sub, my_e);
                else dot = row_dot<V, true>(P.values + a0, P.positions + a0, P.mult, my_s + sub, my_e);
            } else if (win[2 * s + 1]) dot = row_dot<V, false>(vals_s + (size_t)s * cap, cols_s + (size_t)s * cap, P.mult, my_s + sub, my_e);
            else dot = row_dot<V, false>(P.values + a0, P.positions + a0, P.mult, my_s + sub, my_e);
            if (V > 1) {
#pragma unroll
                for (int o = V / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            }
            if (sub == 0 && row < P.rows) write(row, dot);        // the store depends on every shared-memory read above
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);                // this warp is done with the slot
            my_s = nx_s; my_e = nx_e;
            if (++s == stages) { s = 0; phase ^= 1u; }
        }
        ring_s = s; ring_k = phase;
    }
}

// HALO: multi-GPU SpMV whose boundary row groups wait for the peers' halo pushes (compiled out of the single-GPU kernel)
template <int V, bool PLAIN, bool HALO>   // lanes per row; PLAIN: op == ASSIGN and no extra copies of the result
__global__ void __launch_bounds__(ROWS_THREADS) spmv_rows_kernel(const SpmvParams P, const int cap, const int nchunks, const int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: vals[stages][cap] | cols[stages][cap] | full[4] | empty[4] | win[4][2]
    float* vals_s = reinterpret_cast<float*>(smem_raw);
    int* cols_s = reinterpret_cast<int*>(smem_raw + (size_t)stages * cap * sizeof(float));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)stages * cap * 8);
    uint64_t* empty = full + ROWS_MAX_STAGES;
    int* win = reinterpret_cast<int*>(empty + ROWS_MAX_STAGES);   // [stage][2]: window start a0, staged flag
    __shared__ float red_sh[96];
    __shared__ int sh_flag;

    const int tid = threadIdx.x;
    typename std::conditional<PLAIN, PlainWriter, RowWriter>::type write(P);

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ROWS_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // everything above is independent of the previous kernel of the chain; the done flag, the operand vector and (for the
    // multi-GPU case) the exchange counter are not
    smm_pdl_wait();
    if (P.state != nullptr && P.state->done) {
        // an SpMV after the end of the solve: the x update a two-pass CG iteration still owed (VEC_CG_PX) has run by now
        if (blockIdx.x == 0 && threadIdx.x == 0 && P.state->x_owed) P.state->x_owed = 0;
        return;
    }

    int ring_s = 0;
    uint32_t ring_k = 0;
    rows_sweep<V, HALO>(P, cap, nchunks, stages, vals_s, cols_s, full, empty, win, write, (int)blockIdx.x, (int)gridDim.x, ring_s, ring_k);

    smm_pdl_trigger();
    if (P.reduce != RED_NONE) {
        float v[2] = {write.acc0, write.acc1};
        __syncthreads();
        if (grid_sum_last_block<2>(v, P.partials, P.partials_stride, P.ticket, red_sh, &sh_flag)) {
            if (tid == 0) smm_finish(P.finish, P.state, v[0], v[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// cg_persistent_kernel: the whole loop of ConjugateGradient (H:2352-2396) in ONE cooperative launch, for problems that live
// in L2, where an iteration of three kernels is mostly launch ramp-up, tail and the last CTA folding the partial sums
// (1024^2: 25 us per iteration for 90 MB of L2 traffic).  The resident CTAs run the three phases of an iteration --
//   Ap = A p with p.Ap  |  x += alpha p, r -= alpha Ap with r.r  |  p = r + beta p
// -- separated by grid barriers, and every CTA folds the partial sums itself, so the scalars (alpha, beta, the stopping test)
// are known everywhere without a second barrier.
// RESULTS ARE THE SAME BITS AS THE GRAPH DRIVERS': the phases are executed as the VIRTUAL CTAs of the kernels they replace
// (virtual CTA v of spmv_rows_kernel's grid G1 = rows_sweep(first = v, G = G1); virtual CTA v of vec_kernel's grid G2 =
// threads 0..255 with gid = v * 256 + tid), each leaves the partial sum the real CTA would have left, and the fold walks the
// partial sums in the order of grid_sum_last_block (thread t adds partials t, t + T, ...; then the block tree over T threads).
// ---------------------------------------------------------------------------------------------------
struct PersistParams {
    SpmvParams sp;                 // Ap = A p, RED_OUT_AUX with aux = p; partials = reduction slot 0
    int cap, nchunks, stages, G1;  // spmv_rows_kernel's launch configuration
    VecParams xr, pu;              // FCgXR (partials = slot 1) and FCgP
    int G2;                        // vec_kernel's grid
    unsigned int* barrier;         // [0] arrivals, [1] generation (both 0 at launch)
};

// block_sum of smm_internal.cuh for a (virtual) CTA of nw warps inside this CTA: threads of warps >= nw pass zeros
template <int NV>
__device__ __forceinline__ void block_sum_nw(float (&v)[NV], float* sh, const int nw) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sh[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float t = lane < nw ? sh[i * 32 + lane] : 0.0f;
            v[i] = warp_sum(t);
        }
    }
}

__device__ __forceinline__ void grid_barrier(unsigned int* bar, const unsigned int nblocks, unsigned int& generation) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == nblocks - 1u) {
            bar[0] = 0u;                                       // everybody has arrived: nobody touches the count until released
            __threadfence();
            atomicExch(&bar[1], generation + 1u);
        } else {
            unsigned int g;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(bar + 1) : "memory"); } while (g == generation);
        }
        __threadfence();
    }
    ++generation;
    __syncthreads();
}

// the element-wise phase of virtual CTA vb of a grid of G2 CTAs x 256 threads (the body of vec_kernel<F, VEC4 = true, false>)
template <class F>
__device__ __forceinline__ void vec_phase(const VecParams& P, const Scal sc, const int vb, const int G2, float (&red)[2]) {
    if (threadIdx.x >= VEC_THREADS) return;
    const long long stride = (long long)G2 * VEC_THREADS;
    const long long gid = (long long)vb * VEC_THREADS + threadIdx.x;
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
    }
    const long long t = (n4 << 2) + gid;                       // tail (n % 4 elements) by the first threads of virtual CTA 0
    if (gid < 4 && t < P.n) {
        float ein[F::NIN], eout[F::NOUT > 0 ? F::NOUT : 1];
#pragma unroll
        for (int k = 0; k < F::NIN; ++k) ein[k] = P.in[k][t];
        F::apply(sc, ein, eout, red);
#pragma unroll
        for (int k = 0; k < F::NOUT; ++k) P.out[k][t] = eout[k];
    }
}

// totals of `count` partial sums, walked like the last CTA of a kernel of T threads does (grid_sum_last_block); in thread 0
__device__ __forceinline__ void fold_partials(const float* partials, const size_t stride, const int count, const int T, float (&tot)[2], float* sh) {
    float acc[2] = {0.0f, 0.0f};
    if ((int)threadIdx.x < T) {
        for (int b = threadIdx.x; b < count; b += T) {
            acc[0] += __ldcg(&partials[b]);
            acc[1] += __ldcg(&partials[stride + b]);
        }
    }
    __syncthreads();                                           // sh is reused
    block_sum_nw<2>(acc, sh, T / 32);
    tot[0] = acc[0]; tot[1] = acc[1];
}

template <int V>
__global__ void __launch_bounds__(ROWS_THREADS, 5) cg_persistent_kernel(const PersistParams A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* vals_s = reinterpret_cast<float*>(smem_raw);
    int* cols_s = reinterpret_cast<int*>(smem_raw + (size_t)A.stages * A.cap * sizeof(float));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)A.stages * A.cap * 8);
    uint64_t* empty = full + ROWS_MAX_STAGES;
    int* win = reinterpret_cast<int*>(empty + ROWS_MAX_STAGES);
    __shared__ float red_sh[96];
    __shared__ SolveState st;                                  // this CTA's copy of the scalar state: every CTA computes the same one
    const int tid = threadIdx.x;
    const int Gp = gridDim.x;
    SolveState* const global_state = A.sp.state;

    if (tid == 0) {
        for (int s = 0; s < A.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], ROWS_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        st = *global_state;
        if (blockIdx.x != 0) { st.history = nullptr; st.history_cap = 0; }   // one writer for the residual history
    }
    __syncthreads();
    int ring_s = 0;
    uint32_t ring_k = 0;
    unsigned int generation = 0;

    while (!st.done) {
        // ---- Ap = A p, p.Ap -> alpha (H:2353-2358) ----
        for (int v = blockIdx.x; v < A.G1; v += Gp) {
            PlainWriter write(A.sp);
            rows_sweep<V, false>(A.sp, A.cap, A.nchunks, A.stages, vals_s, cols_s, full, empty, win, write, v, A.G1, ring_s, ring_k);
            float part[2] = {write.acc0, write.acc1};
            __syncthreads();
            block_sum_nw<2>(part, red_sh, ROWS_THREADS / 32);
            if (tid == 0) { A.sp.partials[v] = part[0]; A.sp.partials[A.sp.partials_stride + v] = part[1]; }
            __syncthreads();
        }
        grid_barrier(A.barrier, Gp, generation);
        {
            float tot[2];
            fold_partials(A.sp.partials, A.sp.partials_stride, A.G1, ROWS_THREADS, tot, red_sh);
            if (tid == 0) smm_finish(FIN_CG_ALPHA, &st, tot[0], tot[1]);
            __syncthreads();
        }
        // ---- x = fma(alpha, p, x); r = fma(-alpha, Ap, r); r.r -> beta, stopping test (H:2363-2382) ----
        {
            const Scal sc = FCgXR::scal(&st);
            for (int v = blockIdx.x; v < A.G2; v += Gp) {
                float red[2] = {0.0f, 0.0f};
                vec_phase<FCgXR>(A.xr, sc, v, A.G2, red);
                block_sum_nw<2>(red, red_sh, VEC_THREADS / 32);
                if (tid == 0) { A.xr.partials[v] = red[0]; A.xr.partials[A.xr.partials_stride + v] = red[1]; }
                __syncthreads();
            }
        }
        grid_barrier(A.barrier, Gp, generation);
        {
            float tot[2];
            fold_partials(A.xr.partials, A.xr.partials_stride, A.G2, VEC_THREADS, tot, red_sh);
            if (tid == 0) smm_finish(FIN_CG_UPDATE, &st, tot[0], tot[1]);
            __syncthreads();
        }
        if (st.done) break;                                    // the p update of the last iteration is a no-op in the graph drivers too
        // ---- p = fma(beta, p, r) (H:2385-2393) ----
        {
            const Scal sc = FCgP::scal(&st);
            float none[2] = {0.0f, 0.0f};
            for (int v = blockIdx.x; v < A.G2; v += Gp) vec_phase<FCgP>(A.pu, sc, v, A.G2, none);
        }
        grid_barrier(A.barrier, Gp, generation);
    }
    if (blockIdx.x == 0 && tid == 0) {
        global_state->done = st.done; global_state->status = st.status; global_state->iterations = st.iterations;
        global_state->residual = st.residual; global_state->rr = st.rr; global_state->denom = st.denom;
        global_state->alpha = st.alpha; global_state->beta = st.beta;
    }
}

__global__ void row_stats_kernel(const int32_t* __restrict__ start, int rows, int* __restrict__ max_len) {
    int m = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) m = max(m, start[r + 1] - start[r]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_len, m);
}

__global__ void block_row_kernel(const int32_t* __restrict__ start, int rows, int num_blocks, int32_t* __restrict__ block_row) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > num_blocks) return;
    if (q == num_blocks) { block_row[q] = rows; return; }
    const long long target = (long long)q * SPMV_CHUNK;
    int lo = 0, hi = rows;                                    // first r in [0,rows) with start[r] >= target
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if ((long long)start[mid] < target) lo = mid + 1; else hi = mid;
    }
    block_row[q] = lo;
}

// row_perm: the rows of chunk q (one CTA each) ordered by length, longest first, ties in index order -- a rank by counting
// (chunks hold ~ SPMV_CHUNK / mean-row-length rows; chunks of more than SPMV_PERM_MAX_ROWS rows, i.e. mostly empty rows,
// keep the index order and the kernels do not read row_perm for them)
__global__ void __launch_bounds__(256) row_perm_kernel(const int32_t* __restrict__ start, const int32_t* __restrict__ block_row, int32_t* __restrict__ row_perm) {
    __shared__ int len[SPMV_PERM_MAX_ROWS];
    const int r0 = block_row[blockIdx.x], r1 = block_row[blockIdx.x + 1];
    const int n = r1 - r0;
    if (n <= 0) return;
    if (n > SPMV_PERM_MAX_ROWS) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) row_perm[r0 + i] = r0 + i;
        return;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) len[i] = start[r0 + i + 1] - start[r0 + i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int mine = len[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += (len[j] > mine) || (len[j] == mine && j < i);
        row_perm[r0 + rank] = r0 + i;
    }
}

// rows kernel: row starts of every group of R rows relative to the group's staging window (a0 = first entry rounded down to 16
// bytes), R + 1 entries per group so that a row finds its end next to its start; rows past the end get the group's end (empty)
__global__ void start16_kernel(const int32_t* __restrict__ start, const int rows, const int R, const long long total, uint16_t* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long g = t / (R + 1);
    const int i = (int)(t - g * (R + 1));
    const long long r0 = g * R, r = r0 + i;
    const int a0 = start[r0 < rows ? r0 : rows] & ~3;
    out[t] = (uint16_t)(start[r < rows ? r : rows] - a0);
}

__global__ void first_active_kernel(const int32_t* __restrict__ start, int rows, int* out) {
    // firstActiveStart (H:1622-1628): first row i with start[i+1] != 0, or rows
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows && start[i + 1] != 0 && start[i] == 0) atomicMin(out, i);
}

}  // namespace

int smm_csr_analyse(smm_csr* m, cudaStream_t s) {
    // row-length statistics -> kernel choice
    {
        int* d = nullptr;
        SMM_CUDA(cudaMalloc(&d, sizeof(int)));
        SMM_CUDA(cudaMemsetAsync(d, 0, sizeof(int), s));
        if (m->rows > 0) {
            row_stats_kernel<<<1184, 256, 0, s>>>(m->start, m->rows, d);
            SMM_COUNT_LAUNCH(1);
        }
        SMM_CUDA(cudaMemcpyAsync(&m->max_row_len, d, sizeof(int), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
        cudaFree(d);
        const double mean = m->rows ? (double)m->nnz / m->rows : 0.0;
        // lanes per row (V) so that a group of 256/V rows fits a staging window of at most SPMV_ROWS_CAP entries; the
        // window itself is then sized to the matrix (mean group size + 3 % + alignment slack): a smaller window
        // means more resident CTAs per SM, which is what the kernel's latency hiding lives on
        int v = 1;
        while (v < 8 && mean * (SPMV_THREADS / v) > 0.9 * SPMV_ROWS_CAP) v *= 2;
        const bool regular = m->rows >= 4096 && m->max_row_len <= 64 && m->max_row_len <= 4 * mean + 8 &&
                             mean * (SPMV_THREADS / v) <= 0.9 * SPMV_ROWS_CAP;
        m->rows_kernel_lanes = regular ? v : 0;
        int cap = ((int)(mean * (SPMV_THREADS / v) * 1.03) + 8 + 63) & ~63;
        if (cap < 512) cap = 512;
        if (cap > SPMV_ROWS_CAP) cap = SPMV_ROWS_CAP;
        m->rows_kernel_cap = cap;
        const char* env = getenv("SMM_B200_SPMV_KERNEL");           // benchmarking override: "stage" | "rows"
        if (env && !strcmp(env, "stage")) m->rows_kernel_lanes = 0;
        if (env && !strcmp(env, "rows") && m->rows_kernel_lanes == 0) m->rows_kernel_lanes = v;
        // start16 holds row starts relative to their group's window in 16 bits
        // (a group's entries + the up to 3 entries between the 16-byte window start and its first entry must stay below 2^16)
        if (m->rows_kernel_lanes > 0 && (long long)m->max_row_len * (SPMV_THREADS / m->rows_kernel_lanes) + 3 >= 65536) m->rows_kernel_lanes = 0;
        cudaDeviceProp prop;
        SMM_CUDA(cudaGetDeviceProperties(&prop, m->device));
        m->sm_count = prop.multiProcessorCount;
    }
    m->num_blocks = (int)(m->nnz / SPMV_CHUNK) + 1;
    if (m->block_row) { cudaFree(m->block_row); m->block_row = nullptr; }
    SMM_CUDA(cudaMalloc(&m->block_row, sizeof(int32_t) * (size_t)(m->num_blocks + 1)));
    const int threads = 256;
    block_row_kernel<<<(m->num_blocks + 1 + threads - 1) / threads, threads, 0, s>>>(m->start, m->rows, m->num_blocks, m->block_row);
    SMM_COUNT_LAUNCH(1);
    if (m->row_perm) { cudaFree(m->row_perm); m->row_perm = nullptr; }
    static const bool perm_on = [] { const char* e = getenv("SMM_B200_SPMV_PERM"); return !e || atoi(e) != 0; }();
    if (m->rows_kernel_lanes != 1 && m->rows > 0 && perm_on) {      // the product-staging kernels run for this matrix (always, or in exact mode)
        SMM_CUDA(cudaMalloc(&m->row_perm, sizeof(int32_t) * (size_t)m->rows));
        row_perm_kernel<<<m->num_blocks, 256, 0, s>>>(m->start, m->block_row, m->row_perm);
        SMM_COUNT_LAUNCH(1);
    }
    if (m->start16) { cudaFree(m->start16); m->start16 = nullptr; }
    if (m->rows_kernel_lanes > 0) {
        // a regular matrix has rows of at most 64 entries: a group of up to 256 rows spans less than 2^16 entries
        const int R = SPMV_THREADS / m->rows_kernel_lanes;
        const long long total = (long long)((m->rows + R - 1) / R) * (R + 1);
        SMM_CUDA(cudaMalloc(&m->start16, sizeof(uint16_t) * (size_t)(total > 0 ? total : 1)));
        start16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(m->start, m->rows, R, total, m->start16);
        SMM_COUNT_LAUNCH(1);
    }
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

int smm_first_active_start(const smm_csr* m, int* out_host, cudaStream_t s) {
    int* d = nullptr;
    SMM_CUDA(cudaMalloc(&d, sizeof(int)));
    const int rows = m->rows;
    SMM_CUDA(cudaMemcpyAsync(d, &rows, sizeof(int), cudaMemcpyHostToDevice, s));
    if (rows > 0) {
        first_active_kernel<<<(rows + 255) / 256, 256, 0, s>>>(m->start, rows, d);
        SMM_COUNT_LAUNCH(1);
    }
    SMM_CUDA(cudaMemcpyAsync(out_host, d, sizeof(int), cudaMemcpyDeviceToHost, s));
    SMM_CUDA(cudaStreamSynchronize(s));
    cudaFree(d);
    return SMM_OK;
}

// lanes per row of the TMA rows kernel for this matrix and mode, 0 when the product-staging kernel runs instead
int smm_spmv_rows_lanes(const smm_csr* m, int exact) { return (exact && m->rows_kernel_lanes > 1) ? 0 : m->rows_kernel_lanes; }


// ---------------------------------------------------------------------------------------------------
// persistent CG iteration: configuration shared with smm_launch_spmv / smm_launch_vec (the virtual grids must be theirs)
// ---------------------------------------------------------------------------------------------------
namespace {
struct RowsConfig { int cap, stages, nchunks, grid; size_t smem; };
RowsConfig rows_config(const smm_csr* m, const int V) {
    static const int stages_env = [] { const char* e = getenv("SMM_B200_ROWS_STAGES"); return e ? atoi(e) : 0; }();
    static const int cap_env = [] { const char* e = getenv("SMM_B200_ROWS_CAP"); return e ? (atoi(e) & ~3) : 0; }();
    RowsConfig c;
    c.cap = cap_env ? cap_env : m->rows_kernel_cap;
    c.stages = stages_env ? stages_env : (int)(((227 * 1024) / 5 - 1024 - ROWS_MAX_STAGES * 24) / ((size_t)c.cap * 8));
    if (c.stages < 2) c.stages = 2;
    if (c.stages > ROWS_MAX_STAGES) c.stages = ROWS_MAX_STAGES;
    c.smem = (size_t)c.stages * c.cap * 8 + ROWS_MAX_STAGES * 8 * 2 + ROWS_MAX_STAGES * 8;
    const int R = ROWS_CONSUMERS / V;
    c.nchunks = (m->rows + R - 1) / R;
    int per_sm = (int)((227 * 1024) / (c.smem + 1024));
    if (per_sm > 2048 / ROWS_THREADS) per_sm = 2048 / ROWS_THREADS;
    c.grid = m->sm_count * per_sm;
    if (c.grid > c.nchunks) c.grid = c.nchunks;
    return c;
}
}  // namespace

// the persistent iteration applies to matrices of the rows kernel with one lane per row, 16-byte aligned vectors
bool smm_cg_persistent_fits(const smm_csr* m, const float* x, const float* r, const float* p, const float* ap) {
    if (m->rows_kernel_lanes != 1 || m->rows < 4096 || !m->ws) return false;
    return ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(ap)) & 15) == 0;
}

int smm_launch_cg_persistent(const smm_csr* m, SolveState* state, float* x, float* r, float* p, float* ap, cudaStream_t s) {
    smm_workspace* ws = m->ws;
    const RowsConfig rc = rows_config(m, 1);
    PersistParams A;
    SpmvParams& P = A.sp;
    P.start = m->start; P.positions = m->positions; P.values = m->values; P.block_row = m->block_row; P.row_perm = m->row_perm; P.start16 = m->start16;
    P.rows = m->rows; P.nnz = (int)m->nnz; P.nnz_alloc = (int)m->nnz_alloc; P.op = SMM_OP_ASSIGN; P.exact = 0;
    P.lhs = nullptr; P.mult = p; P.out = ap; P.copy1 = P.copy2 = P.copy3 = nullptr;
    P.aux = p; P.reduce = RED_OUT_AUX; P.finish = FIN_CG_ALPHA; P.state = state;
    P.partials = ws->partials; P.partials_stride = ws->partials_cap; P.ticket = nullptr; P.halo = nullptr;
    A.cap = rc.cap; A.nchunks = rc.nchunks; A.stages = rc.stages; A.G1 = rc.grid;
    A.G2 = smm_vec_grid(ws, m->rows, true);
    if (ws->partials_cap < (size_t)A.G1 || ws->partials_cap < (size_t)A.G2) { smm_set_error("persistent CG: reduction workspace too small"); return SMM_E_STATE; }
    auto vec = [&](VecParams& V) {
        V.n = m->rows; V.state = state; V.finish = FIN_NONE; V.ticket = nullptr; V.halo = nullptr;
        for (int k = 0; k < 5; ++k) V.in[k] = nullptr;
        for (int k = 0; k < 3; ++k) V.out[k] = nullptr;
        V.partials = ws->partials + (size_t)1 * 2 * ws->partials_cap;     // reduction slot 1, as in the graph drivers
        V.partials_stride = ws->partials_cap;
    };
    vec(A.xr); A.xr.in[0] = x; A.xr.in[1] = p; A.xr.in[2] = r; A.xr.in[3] = ap; A.xr.out[0] = x; A.xr.out[1] = r;
    vec(A.pu); A.pu.in[0] = p; A.pu.in[1] = r; A.pu.out[0] = p;
    A.barrier = ws->grid_barrier;
    static int per_sm_dev[SMM_MAX_DEVICES] = {0};
    static size_t attr_dev[SMM_MAX_DEVICES] = {0};
    int per_sm;
    {
        std::lock_guard<std::mutex> lk(g_smm_attr_mu);
        const int d = m->device % SMM_MAX_DEVICES;
        if (attr_dev[d] != rc.smem) {
            SMM_CUDA(cudaFuncSetAttribute(cg_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rc.smem));
            int n = 0;
            SMM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, cg_persistent_kernel<1>, ROWS_THREADS, rc.smem));
            per_sm_dev[d] = n;
            attr_dev[d] = rc.smem;
        }
        per_sm = per_sm_dev[d];
    }
    if (per_sm < 1) { smm_set_error("persistent CG: the kernel does not fit an SM"); return SMM_E_STATE; }
    int grid = per_sm * m->sm_count;                           // every CTA resident (cooperative launch): the grid barrier relies on it
    const int most = A.G1 > A.G2 ? A.G1 : A.G2;
    if (grid > most) grid = most;
    SMM_CUDA(cudaMemsetAsync(ws->grid_barrier, 0, 2 * sizeof(unsigned int), s));
    void* args[] = {&A};
    SMM_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(cg_persistent_kernel<1>), dim3(grid), dim3(ROWS_THREADS), args, rc.smem, s));
    SMM_COUNT_LAUNCH(1);
    return SMM_OK;
}

int smm_launch_spmv(const SpmvArgs& a, cudaStream_t s) {
    const smm_csr* m = a.m;
    if (m->rows == 0) return SMM_OK;
    SpmvParams P;
    P.start = m->start; P.positions = m->positions; P.values = m->values; P.block_row = m->block_row; P.row_perm = m->row_perm; P.start16 = m->start16;
    P.rows = m->rows; P.nnz = (int)m->nnz; P.nnz_alloc = (int)m->nnz_alloc; P.op = a.op; P.exact = a.exact;
    P.lhs = a.lhs; P.mult = a.mult; P.out = a.out;
    P.copy1 = a.copy1; P.copy2 = a.copy2; P.copy3 = a.copy3;
    P.aux = a.aux; P.reduce = a.reduce; P.finish = a.finish; P.state = a.state;
    P.partials = nullptr; P.partials_stride = 0; P.ticket = nullptr;
    P.halo = nullptr;
    if (a.reduce != RED_NONE) {
        smm_workspace* ws = m->ws;
        if (!ws || ws->partials_cap < (size_t)m->num_blocks) { smm_set_error("spmv: reduction workspace too small"); return SMM_E_STATE; }
        P.partials = ws->partials + (size_t)a.slot * 2 * ws->partials_cap;
        P.partials_stride = ws->partials_cap;
        P.ticket = ws->tickets + a.slot;
    }
    const int V = smm_spmv_rows_lanes(m, a.exact);             // exact mode needs one lane per row
    cudaError_t le = cudaSuccess;
    if (a.halo_wait != nullptr && V <= 0) { smm_set_error("spmv: the fused halo wait needs the rows kernel"); return SMM_E_STATE; }
    if (V > 0) {
        P.halo = static_cast<const HaloWaitDev*>(a.halo_wait);
        struct RowsEnv { int stages, cap; };
        static const RowsEnv env = [] {                                  // read once (thread-safe initialisation)
            const char* e1 = getenv("SMM_B200_ROWS_STAGES");
            const char* e2 = getenv("SMM_B200_ROWS_CAP");
            return RowsEnv{e1 ? atoi(e1) : 0,                            // 0: as many stages as five resident CTAs per SM allow
                           e2 ? (atoi(e2) & ~3) : 0};                    // 0: per-matrix window from the analysis
        }();
        const int stages_env = env.stages, cap_env = env.cap;
        const int cap = cap_env ? cap_env : m->rows_kernel_cap;
        // Five resident CTAs per SM hide the gather latency best (measured: 7-point stencil 3 stages x 14.8 KB, 5 CTAs 1.46 ms
        // vs 3 CTAs with 4 stages 1.97 ms; 5-point stencil 4 stages x 10.8 KB, 5 CTAs 14.4 us vs 6 CTAs with 3 stages 18.4 us on
        // 1024^2): give each CTA a fifth of the shared memory and turn all of it into pipeline depth.
        int stages = stages_env ? stages_env : (int)(((227 * 1024) / 5 - 1024 - ROWS_MAX_STAGES * 24) / ((size_t)cap * 8));
        if (stages < 2) stages = 2;
        if (stages > ROWS_MAX_STAGES) stages = ROWS_MAX_STAGES;
        const size_t smem = (size_t)stages * cap * 8 + ROWS_MAX_STAGES * 8 * 2 + ROWS_MAX_STAGES * 8;
        const int R = ROWS_CONSUMERS / V;
        const int nchunks = (m->rows + R - 1) / R;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 2048 / ROWS_THREADS) per_sm = 2048 / ROWS_THREADS;
        int grid = m->sm_count * per_sm;
        if (grid > nchunks) grid = nchunks;
        if (a.reduce != RED_NONE && m->ws->partials_cap < (size_t)grid) { smm_set_error("spmv: reduction workspace too small"); return SMM_E_STATE; }
        static size_t attr_smem[SMM_MAX_DEVICES] = {0};                  // function attributes are per device
        std::lock_guard<std::mutex> attr_lock(g_smm_attr_mu);
        if (attr_smem[m->device % SMM_MAX_DEVICES] != smem) {
#define SMM_ROWS_ATTR(V_, PL_) SMM_CUDA(cudaFuncSetAttribute(spmv_rows_kernel<V_, PL_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    SMM_CUDA(cudaFuncSetAttribute(spmv_rows_kernel<V_, PL_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
            SMM_ROWS_ATTR(1, true); SMM_ROWS_ATTR(2, true); SMM_ROWS_ATTR(4, true); SMM_ROWS_ATTR(8, true);
            SMM_ROWS_ATTR(1, false); SMM_ROWS_ATTR(2, false); SMM_ROWS_ATTR(4, false); SMM_ROWS_ATTR(8, false);
#undef SMM_ROWS_ATTR
            attr_smem[m->device % SMM_MAX_DEVICES] = smem;
        }
        const bool plain = a.op == SMM_OP_ASSIGN && !a.copy1 && !a.copy2 && !a.copy3;
#define SMM_ROWS_LAUNCH(V_) \
    if (P.halo != nullptr) { \
        if (plain) le = smm_launch_chain(spmv_rows_kernel<V_, true, true>, grid, ROWS_THREADS, smem, s, P, cap, nchunks, stages); \
        else le = smm_launch_chain(spmv_rows_kernel<V_, false, true>, grid, ROWS_THREADS, smem, s, P, cap, nchunks, stages); \
    } else if (plain) le = smm_launch_chain(spmv_rows_kernel<V_, true, false>, grid, ROWS_THREADS, smem, s, P, cap, nchunks, stages); \
    else le = smm_launch_chain(spmv_rows_kernel<V_, false, false>, grid, ROWS_THREADS, smem, s, P, cap, nchunks, stages)
        switch (V) {
            case 1: SMM_ROWS_LAUNCH(1); break;
            case 2: SMM_ROWS_LAUNCH(2); break;
            case 4: SMM_ROWS_LAUNCH(4); break;
            default: SMM_ROWS_LAUNCH(8); break;
        }
#undef SMM_ROWS_LAUNCH
    } else {
        // knobs for measurements: SMM_B200_SPMV_DEPTH = 2 | 4 (vectors in flight per thread), SMM_B200_SPMV_HINTS = 0 | 1 (L2 eviction priorities)
        static const int depth = [] { const char* e = getenv("SMM_B200_SPMV_DEPTH"); const int v = e ? atoi(e) : 2; return v == 4 ? 4 : 2; }();
        static const bool hints = [] { const char* e = getenv("SMM_B200_SPMV_HINTS"); return e ? atoi(e) != 0 : false; }();
        if (depth == 4) { if (hints) le = smm_launch_chain(spmv_kernel<4, true>, m->num_blocks, SPMV_THREADS, 0, s, P); else le = smm_launch_chain(spmv_kernel<4, false>, m->num_blocks, SPMV_THREADS, 0, s, P); }
        else { if (hints) le = smm_launch_chain(spmv_kernel<2, true>, m->num_blocks, SPMV_THREADS, 0, s, P); else le = smm_launch_chain(spmv_kernel<2, false>, m->num_blocks, SPMV_THREADS, 0, s, P); }
    }
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(le);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}
