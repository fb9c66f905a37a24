// sgs_tiles_setup.cu -- getPreconditioner()'s set-up for the tile-level sweep schedule, on the device.
//
// The reference's SGSPreconditioner has no set-up at all (H:1172-1186): apply() walks A's rows in place.  The tile schedule
// of sgs_tiles.cu needs a layout (tiles, tile levels, rows inside a tile by internal level, operand positions, push lists),
// and building it on the host meant downloading start / positions (0.26 s at 256^3), 0.3 s of host threads and 0.23 s of
// uploads for arrays that are only ever read by kernels.  Here the same arrays are produced by kernels from the CSR arrays
// where they lie in HBM -- nothing is downloaded but a handful of flags and counts.
//
// It is the SAME layout, array for array and bit for bit, as layout_sweep() of sgs_tiles.cu in its default mode (tiles in
// tile-level order, single-tile chains): the tests compare fingerprints of every array the sweep kernels read between a
// handle built here and one built on the host (tests/test_gpu_parity.py::test_tile_layout_device_equals_host), and the host
// code stays pinned by tests/test_tile_layout_cpu.py.  The steps mirror the host's:
//   1. proposal: distinct |col - row| over the head / middle / tail rows -> grid shape -> geometric tile of every row;
//   2. rows of every tile, ascending (count, exclusive scan, scatter, a warp sorts its tile's <= 64 rows by rank);
//   3. per sweep: distinct predecessor tiles (a thread per tile, in the order the rows meet them), tile levels = longest
//      path, by relaxation until nothing changes (a cyclic tile graph never settles: bounded, then left to the host path to
//      refuse), tiles ordered by level, stable in the tile id (per-block level histograms, a column scan, ranks inside a block);
//   4. per sweep and tile (a thread per tile, the host's loops verbatim): internal levels, rows by internal level, operand
//      positions in the reference's operand order, in-tile push lists.
// Every "does not verify" of the host code (a tile with more than 64 rows, more than 8 predecessor tiles, 64 internal levels,
// a row with more than three consumers inside its tile) raises a flag here and returns false; the caller then runs the host
// path, which decides.  The opt-in chain / cluster layouts are only built on the host.
//
// The IC(0) / ILU(0) factorisations run here too (smm_sgs_factorize_dev, further down): level by level along the forward tile
// schedule, one thread per row doing the host code's operations in the host code's order -- the same factor bits, without
// the download / host factorisation / upload that cost 1.1 s at 256^3.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "sgs_internal.cuh"

namespace {

constexpr int SETUP_T = 256;

struct DevMem {                                                // scratch that goes away when the build returns
    std::vector<void*> ptrs;
    ~DevMem() { for (void* q : ptrs) cudaFree(q); }
    template <class T>
    T* get(size_t n) {
        void* q = nullptr;
        if (cudaMalloc(&q, sizeof(T) * (n ? n : 1)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        ptrs.push_back(q);
        return static_cast<T*>(q);
    }
};

inline unsigned int blocks_for(long long n, int per = SETUP_T) { return (unsigned int)((n + per - 1) / per); }

// ---- diagonals (find_diagonals of sgs.cu: H:1678-1680 empty row, H:1691 missing diagonal) ---------------------------------
__global__ void setup_diag_kernel(int rows, const int32_t* __restrict__ start, const int32_t* __restrict__ pos, int32_t* diag, int* info) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int w = 0;
    bool bad = false;
    if (r < rows) {
        int k = start[r];
        const int b = k, e = start[r + 1];
        if (e == k) bad = true;
        else {
            while (k < e && pos[k] < r) ++k;
            if (k >= e || pos[k] != r) bad = true;
            else { diag[r] = k; w = max(k - b, e - 1 - k); }
        }
    }
    w = __reduce_max_sync(0xFFFFFFFFu, w);
    bad = __any_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (w > 0) atomicMax(&info[1], w);
        if (bad) info[0] = 1;
    }
}

// ---- 1. proposal ---------------------------------------------------------------------------------------------------------
// distinct |col - row| > 0 over rows [0, sample), [rows / 2, rows / 2 + sample), [rows - sample, rows): the sample of
// smm_sgs_detect_grid.  offs[0..2]: the set (0 = empty); offs[3] = 1 when a fourth value turned up.
__global__ void setup_offsets_kernel(int rows, const int32_t* __restrict__ start, const int32_t* __restrict__ pos, int sample, int* offs) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int which = (int)(t / sample);
    const long long o = t % sample;
    const long long r = which == 0 ? o : which == 1 ? (long long)(rows / 2) + o : (long long)rows - sample + o;
    if (which > 2 || r < 0 || r >= rows) return;
    for (int k = start[r]; k < start[r + 1]; ++k) {
        long long dl = (long long)pos[k] - r;
        if (dl < 0) dl = -dl;
        if (dl == 0) continue;
        const int d = (int)dl;
        int i = 0;
        for (; i < 3; ++i) {
            int cur = ((volatile int*)offs)[i];
            if (cur == 0) cur = atomicCAS(&offs[i], 0, d);
            if (cur == 0 || cur == d) break;
        }
        if (i == 3) offs[3] = 1;
    }
}

__global__ void setup_tile_of_row_kernel(int rows, long long nx, long long ny, int ti, int tj, int tk, long long TI, long long TJ, int32_t* cl, int* cnt) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const long long i = r % nx, j = (r / nx) % ny, k = r / (nx * ny);
    const int c = (int)(((k / tk) * TJ + j / tj) * TI + i / ti);
    cl[r] = c;
    atomicAdd(&cnt[c], 1);
}

// ---- exclusive scan of int32 (4096 elements per block, block sums scanned recursively) -------------------------------------
constexpr int SCAN_BLOCK = 4096;
__global__ void __launch_bounds__(1024) scan_block_kernel(const int* in, int* out, long long n, int* sums) {
    __shared__ int wsum[32];
    const long long base = (long long)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
    int v[4], t = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = base + i < n ? in[base + i] : 0; t += v[i]; }
    int inc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((threadIdx.x & 31) >= d) inc += o; }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = wsum[threadIdx.x], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)threadIdx.x >= d) wi += o; }
        wsum[threadIdx.x] = wi - w;                            // exclusive offset of the warp
        if (threadIdx.x == 31 && sums) sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = inc - t + wsum[threadIdx.x >> 5];
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
}
__global__ void scan_add_kernel(int* out, long long n, const int* __restrict__ offs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += offs[i / SCAN_BLOCK];
}
bool exclusive_scan(const int* in, int* out, long long n, DevMem& mem, cudaStream_t s) {
    const long long nb = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
    if (nb <= 1) { scan_block_kernel<<<1, 1024, 0, s>>>(in, out, n, nullptr); return true; }
    int* sums = mem.get<int>((size_t)nb);
    if (!sums) return false;
    scan_block_kernel<<<(unsigned)nb, 1024, 0, s>>>(in, out, n, sums);
    if (!exclusive_scan(sums, sums, nb, mem, s)) return false;
    scan_add_kernel<<<blocks_for(n), SETUP_T, 0, s>>>(out, n, sums);
    return true;
}

// ---- 2. rows of every tile, ascending --------------------------------------------------------------------------------------
__global__ void setup_scatter_kernel(int rows, const int32_t* __restrict__ cl, int* cur, int32_t* crow) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < rows) crow[atomicAdd(&cur[cl[r]], 1)] = (int32_t)r;
}
// a warp per tile: rank of each of its <= 64 rows among them (rows are distinct)
__global__ void setup_sort_rows_kernel(int ncl, const int* __restrict__ cptr, int32_t* crow, int* flags) {
    const int a = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (a >= ncl) return;
    const int b = cptr[a], n = cptr[a + 1] - b;
    if (n > TILE) { if (lane == 0) flags[0] = 1; return; }     // the proposal puts more than 64 rows into a tile
    const int r0 = lane < n ? crow[b + lane] : INT32_MAX, r1 = lane + 32 < n ? crow[b + 32 + lane] : INT32_MAX;
    int k0 = 0, k1 = 0;
    for (int j = 0; j < 32; ++j) {
        const int x0 = __shfl_sync(0xFFFFFFFFu, r0, j), x1 = __shfl_sync(0xFFFFFFFFu, r1, j);
        k0 += (x0 < r0) + (x1 < r0);
        k1 += (x0 < r1) + (x1 < r1);
    }
    __syncwarp();
    if (lane < n) crow[b + k0] = r0;
    if (lane + 32 < n) crow[b + k1] = r1;
}

// ---- 3. tile graph ---------------------------------------------------------------------------------------------------------
struct Sweep {
    bool forward;
    const int32_t* start;
    const int32_t* pos;
    const int32_t* diag;
    __device__ int dep_begin(int r) const { return forward ? start[r] : diag[r] + 1; }
    __device__ int dep_end(int r) const { return forward ? diag[r] : start[r + 1]; }
};

__global__ void setup_preds_kernel(Sweep S, int ncl, const int* __restrict__ cptr, const int32_t* __restrict__ crow, const int32_t* __restrict__ cl,
                                   int32_t* pred, uint8_t* npred, int* flags) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ncl) return;
    int32_t pa[MAX_PREDS];
    int n = 0;
    for (int q = cptr[a]; q < cptr[a + 1]; ++q) {
        const int r = crow[q];
        for (int k = S.dep_begin(r); k < S.dep_end(r); ++k) {
            const int b = cl[S.pos[k]];
            if (b == a) continue;
            int i = 0;
            while (i < n && pa[i] != b) ++i;
            if (i == n) {
                if (n == MAX_PREDS) { flags[0] = 1; return; }
                pa[n++] = b;
            }
        }
    }
    for (int i = 0; i < MAX_PREDS; ++i) pred[(size_t)a * MAX_PREDS + i] = i < n ? pa[i] : -1;
    npred[a] = (uint8_t)n;
}

// longest path by relaxation, in place: levels only grow, a launch that changes nothing has found the fixed point
__global__ void setup_relax_kernel(int ncl, const int32_t* __restrict__ pred, const uint8_t* __restrict__ npred, int* level, int* changed) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ncl) return;
    int l = 0;
    for (int i = 0; i < npred[a]; ++i) l = max(l, ((volatile int*)level)[pred[(size_t)a * MAX_PREDS + i]] + 1);
    if (l > level[a]) { level[a] = l; *changed = 1; }
}
__global__ void setup_max_kernel(int n, const int* __restrict__ v, int* out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    int x = a < n ? v[a] : 0;
    x = __reduce_max_sync(0xFFFFFFFFu, x);
    if ((threadIdx.x & 31) == 0 && x > 0) atomicMax(out, x);
}

// tiles by level, stable in the tile id (forward: ascending; backward: descending).  idx = the tile's place in that id order.
constexpr int HIST_BLOCK = 256;
__global__ void setup_level_hist_kernel(int ncl, bool forward, const int* __restrict__ level, int nlev, int* H) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ncl) return;
    const int a = forward ? idx : ncl - 1 - idx;
    atomicAdd(&H[(size_t)(idx / HIST_BLOCK) * nlev + level[a]], 1);
}
__global__ void setup_level_colscan_kernel(int nb, int nlev, int* H, int* total) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlev) return;
    int run = 0;
    for (int b = 0; b < nb; ++b) { const int v = H[(size_t)b * nlev + l]; H[(size_t)b * nlev + l] = run; run += v; }
    total[l] = run;
}
__global__ void __launch_bounds__(HIST_BLOCK) setup_tile_order_kernel(int ncl, bool forward, const int* __restrict__ level, int nlev, const int* __restrict__ H,
                                                                      const int* __restrict__ lptr, int32_t* tile_of) {
    __shared__ int lev[HIST_BLOCK];
    const int idx = blockIdx.x * HIST_BLOCK + threadIdx.x;
    const int a = forward ? idx : ncl - 1 - idx;
    const int mine = idx < ncl ? level[a] : -1;
    lev[threadIdx.x] = mine;
    __syncthreads();
    if (idx >= ncl) return;
    int rank = 0;
    for (int j = 0; j < (int)threadIdx.x; ++j) rank += lev[j] == mine;
    tile_of[a] = lptr[mine] + H[(size_t)blockIdx.x * nlev + mine] + rank;
}

// ---- 4. inside the tiles -----------------------------------------------------------------------------------------------------
// internal levels in dependency order, rows by internal level (stable in the sweep's row order): order / steps / where
__global__ void setup_steps_kernel(Sweep S, int ncl, long long ntl, const int* __restrict__ cptr, const int32_t* __restrict__ crow, const int32_t* __restrict__ cl,
                                   const int32_t* __restrict__ tile_of, int8_t* ilev, int32_t* order, uint8_t* steps, int32_t* where, int* flags) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ncl) return;
    const int n = cptr[a + 1] - cptr[a];
    const int32_t* R = crow + cptr[a];
    int at[MAX_STEPS + 1];
    for (int l = 0; l <= MAX_STEPS; ++l) at[l] = 0;
    int nl = 0;
    for (int q = 0; q < n; ++q) {
        const int r = S.forward ? R[q] : R[n - 1 - q];
        int l = 0;
        for (int k = S.dep_begin(r); k < S.dep_end(r); ++k) {
            const int c = S.pos[k];
            if (cl[c] == a) l = max(l, (int)((volatile int8_t*)ilev)[c] + 1);
        }
        if (l >= MAX_STEPS) { flags[0] = 1; return; }
        ilev[r] = (int8_t)l;
        at[l + 1]++;
        nl = max(nl, l + 1);
    }
    for (int l = 0; l < nl; ++l) at[l + 1] += at[l];
    const long long t = tile_of[a];
    steps[ntl * TILE + t] = (uint8_t)nl;
    for (int q = 0; q < n; ++q) {
        const int r = S.forward ? R[q] : R[n - 1 - q];
        const int l = ((volatile int8_t*)ilev)[r];
        const int i = at[l]++;
        steps[t * TILE + i] = (uint8_t)l;
        order[t * TILE + i] = r;
        where[r] = (int32_t)(t * TILE + i);
    }
}

// entries [tile][slot][64] in the reference's operand order (ascending columns forward, descending backward) and the push
// lists: where inside the tile's operand staging a row's result has to go
__global__ void setup_entries_kernel(Sweep S, int ncl, int width, const int* __restrict__ cptr, const int32_t* __restrict__ crow, const int32_t* __restrict__ tile_of,
                                     const int32_t* __restrict__ where, int32_t* ecol, int32_t* eidx, uint32_t* push, int* flags) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ncl) return;
    const long long t = tile_of[a];
    uint32_t pu[TILE];
    uint8_t npush[TILE];
    for (int i = 0; i < TILE; ++i) { pu[i] = (uint32_t)(i * TILE_MAX_W) * 0x01010101u; npush[i] = 0; }
    for (int q = cptr[a]; q < cptr[a + 1]; ++q) {
        const int r = crow[q];
        const int i = where[r] & (TILE - 1);
        const int cnt = S.dep_end(r) - S.dep_begin(r);
        for (int e = 0; e < cnt; ++e) {
            const int srci = S.forward ? S.start[r] + e : S.start[r + 1] - 1 - e;
            const long long at = (t * width + e) * TILE + i;
            const int w = where[S.pos[srci]];
            ecol[at] = w;
            eidx[at] = srci;
            if ((w >> 6) == t) {                               // produced inside the tile: the producer pushes it
                const int j = w & (TILE - 1);
                if (npush[j] == 3) { flags[0] = 1; return; }   // a row with more than three consumers inside its tile
                const int sh = 8 * npush[j]++;
                pu[j] = (pu[j] & ~(0xFFu << sh)) | ((uint32_t)(i * TILE_MAX_W + e) << sh);
            }
        }
    }
    for (int i = 0; i < TILE; ++i) push[t * TILE + i] = pu[i];
}

__global__ void setup_ypos_kernel(long long npos, const int32_t* __restrict__ order_b, const int32_t* __restrict__ where_f, int32_t* yp) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < npos) { const int o = order_b[t]; yp[t] = o >= 0 ? where_f[o] : 0; }
}

__global__ void setup_exclusive_small_kernel(int n, const int* __restrict__ in, int* out) {   // n levels: a few hundred
    if (blockIdx.x == 0 && threadIdx.x == 0) { int run = 0; for (int i = 0; i < n; ++i) { const int v = in[i]; out[i] = run; run += v; } }
}

void release_tiles(smm_precond* p) {
    cudaGetLastError();
    cudaFree(p->order_fwd); cudaFree(p->order_bwd); cudaFree(p->ypos); cudaFree(p->yperm); cudaFree(p->xperm);
    p->order_fwd = p->order_bwd = p->ypos = nullptr;
    p->yperm = p->xperm = nullptr;
    for (int w = 0; w < 2; ++w) {
        cudaFree(p->ecol[w]); cudaFree(p->eidx[w]); cudaFree(p->tile_steps[w]); cudaFree(p->tile_push[w]); cudaFree(p->eval[w]); cudaFree(p->dval[w]);
        p->ecol[w] = p->eidx[w] = nullptr; p->tile_steps[w] = nullptr; p->tile_push[w] = nullptr; p->eval[w] = p->dval[w] = nullptr;
        p->esize[w] = 0;
    }
    p->threads_fwd = p->threads_bwd = 0;
    p->tile_level_ptr.clear();
}

// ---- IC(0) / ILU(0) factorisation in the order of the forward tile schedule ------------------------------------------------
// A row's factor entries depend on the rows its strict lower triangle points to -- the forward sweep's dependencies -- so the
// factorisation runs tile level by tile level (one launch per level, a warp per tile, the tile's rows step by step), every row
// by ONE thread that performs the host code's operations in the host code's order (sgs.cu: ic0_factorize_host /
// ilu0_factorize_host, which restate H:1839-1928 / H:1723-1790): the factor has the same bits.
struct FactorArgs {
    const int32_t* order;        // forward sweep: [tiles * 64] row or -1
    const uint8_t* steps;        // [tiles * 64] step of the row, then [tiles] steps of the tile
    long long ntl;
    const int32_t* start;
    const int32_t* pos;
    const int32_t* diag;
    const float* a;
    float* f;                    // factor, A's pattern
    float* dinv;                 // [rows] reciprocal pivots
    int* flags;                  // [0] ILU(0): pivot not > 1e-6 ; [1] IC(0): non-finite pivot (left to the host code)
};

__device__ __forceinline__ void ic0_row(const FactorArgs& A, const int j) {
    const int rs = A.start[j], dj = A.diag[j];
    float* l = A.f;
    for (int e = rs; e < dj; ++e) {
        const int i = A.pos[e];                                // entry (j, i), i < j
        float sum = 0.0f;
        int ki = A.start[i];
        const int di = A.diag[i];
        for (int ek = rs; ek < e; ++ek) {                      // row j's columns k < i, ascending (H:1900-1907)
            const int k = A.pos[ek];
            while (ki < di && A.pos[ki] < k) ++ki;
            if (ki < di && A.pos[ki] == k) sum = __fadd_rn(sum, __fmul_rn(l[ki], l[ek]));
        }
        l[e] = __fmul_rn(__fsub_rn(A.a[e], sum), A.dinv[i]);   // H:1914
    }
    float dsum = 0.0f;
    for (int e = rs; e < dj; ++e) dsum = __fadd_rn(dsum, __fmul_rn(l[e], l[e]));   // H:1868-1872
    const float d = __fsqrt_rn(__fsub_rn(A.a[dj], dsum));      // H:1879
    l[dj] = d;
    A.dinv[j] = __fdiv_rn(1.0f, d);                            // H:1883
    if (!(fabsf(d) <= 3.4028234e38f)) A.flags[1] = 1;          // NaN / inf: the sign and payload of a host NaN are not reproduced here
}

__device__ __forceinline__ void ilu0_row(const FactorArgs& A, const int row) {
    const int rs = A.start[row], re = A.start[row + 1], dg = A.diag[row];
    float* lu = A.f;
    for (int kp = rs; kp < dg; ++kp) {                         // columns k < row, ascending
        const int k = A.pos[kp];
        const float alpha = __fmul_rn(lu[kp], A.dinv[k]);      // H:1762
        lu[kp] = alpha;
        for (int cp = A.start[k + 1] - 1; cp > A.diag[k]; --cp) {              // row k of U, strictly right of its diagonal
            const int c = A.pos[cp];
            int ci = rs;
            while (ci < re && A.pos[ci] < c) ++ci;             // the host's column_index lookup: columns are ascending and distinct
            if (ci < re && A.pos[ci] == c) lu[ci] = __fsub_rn(lu[ci], __fmul_rn(alpha, lu[cp]));   // H:1766-1768, two roundings
        }
    }
    const float pivot = lu[dg];
    if (!(fabsf(pivot) > 1e-6f)) A.flags[0] = 1;               // H:1774
    A.dinv[row] = __fdiv_rn(1.0f, pivot);                      // H:1778
}

template <int KIND>
__global__ void factor_level_kernel(const FactorArgs A, const int tile0, const int tile1) {
    const int t = tile0 + (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= tile1) return;
    const int nsteps = A.steps[A.ntl * TILE + t];
    int row[2], step[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) { row[k] = A.order[(long long)t * TILE + k * 32 + lane]; step[k] = A.steps[(long long)t * TILE + k * 32 + lane]; }
    for (int s = 0; s < nsteps; ++s) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (row[k] >= 0 && step[k] == s) { if (KIND == 1) ic0_row(A, row[k]); else ilu0_row(A, row[k]); }
        }
        __threadfence_block();
        __syncwarp();
    }
}

// IC(0): the transpose goes into the upper triangle of the same pattern (H:1916-1917)
__global__ void ic0_transpose_kernel(int rows, const int32_t* __restrict__ start, const int32_t* __restrict__ pos, const int32_t* __restrict__ diag, float* l) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rows) return;
    for (int e = start[j]; e < diag[j]; ++e) {
        const int i = pos[e];
        int lo = diag[i] + 1, hi = start[i + 1];               // first position in row i's upper part with column >= j
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pos[mid] < j) lo = mid + 1; else hi = mid; }
        if (lo < start[i + 1] && pos[lo] == j) l[lo] = l[e];
    }
}

}  // namespace

void smm_sgs_tiles_release(smm_precond* p) {
    release_tiles(p);
    cudaFree(p->tile_push2[0]); cudaFree(p->tile_push2[1]);
    p->tile_push2[0] = p->tile_push2[1] = nullptr;
    p->tiled = false;
    p->tile_blocks = 0;
    p->tile_width = 0;
    p->tile_levels[0] = p->tile_levels[1] = 0;
}

int smm_sgs_factorize_dev(smm_precond* p, int* code) {
    const smm_csr* m = p->m;
    if (!p->tiled || p->lined || p->tile_chain[0] != 1 || p->tile_level_ptr.size() != (size_t)p->tile_levels[0] + 1 || !p->diag_pos || m->nnz <= 0) return SMM_E_STATE;
    if (p->kind != 1 && p->kind != 2) return SMM_E_INVALID;
    cudaStream_t s = smm_default_stream();
    DevMem mem;
    float* dinv = mem.get<float>((size_t)m->rows);
    int* flags = mem.get<int>(2);
    if (!dinv || !flags) return SMM_E_STATE;
    SMM_CUDA(cudaMalloc(&p->factor, sizeof(float) * (size_t)m->nnz));
    auto fail = [&](int rc) { cudaFree(p->factor); p->factor = nullptr; return rc; };
    cudaError_t e = cudaMemsetAsync(flags, 0, 2 * sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(dinv, 0, sizeof(float) * (size_t)m->rows, s);
    if (e == cudaSuccess) e = p->kind == 1 ? cudaMemsetAsync(p->factor, 0, sizeof(float) * (size_t)m->nnz, s)          // l.assign(nnz, 0)
                                           : cudaMemcpyAsync(p->factor, m->values, sizeof(float) * (size_t)m->nnz, cudaMemcpyDeviceToDevice, s);   // lu = a, H:1732
    if (e != cudaSuccess) return fail(smm_cuda_fail(e, "factorisation set-up", __FILE__, __LINE__));
    const FactorArgs A{p->order_fwd, p->tile_steps[0], p->threads_fwd / TILE, m->start, m->positions, p->diag_pos, m->values, p->factor, dinv, flags};
    for (int l = 0; l < p->tile_levels[0]; ++l) {
        const int t0 = p->tile_level_ptr[(size_t)l], t1 = p->tile_level_ptr[(size_t)l + 1];
        if (t1 <= t0) continue;
        const unsigned int grid = blocks_for(32ll * (t1 - t0), 128);
        if (p->kind == 1) factor_level_kernel<1><<<grid, 128, 0, s>>>(A, t0, t1);
        else factor_level_kernel<2><<<grid, 128, 0, s>>>(A, t0, t1);
    }
    if (p->kind == 1) ic0_transpose_kernel<<<blocks_for(m->rows), SETUP_T, 0, s>>>(m->rows, m->start, m->positions, p->diag_pos, p->factor);
    int h[2] = {0, 0};
    e = cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(smm_cuda_fail(e, "factorisation", __FILE__, __LINE__));
    if (h[1] || h[0]) return fail(SMM_E_STATE);                 // a failed factorisation is reproduced by the host code (which rows it still wrote, its NaN bits)
    *code = 0;
    return SMM_OK;
}


int smm_sgs_diagonals_dev(const smm_csr* m, int32_t* diag_dev, bool* valid, int* width) {
    *valid = false;
    *width = 0;
    if (!(m->first_active_start == 0 || m->rows == 0)) return SMM_OK;            // H:1668-1670
    if (m->rows == 0) { *valid = true; return SMM_OK; }
    cudaStream_t s = smm_default_stream();
    int* info = nullptr;
    SMM_CUDA(cudaMalloc(&info, 2 * sizeof(int)));
    int h[2] = {0, 0};
    cudaError_t e = cudaMemsetAsync(info, 0, 2 * sizeof(int), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(diag_dev, 0, sizeof(int32_t) * (size_t)m->rows, s);
    if (e == cudaSuccess) {
        setup_diag_kernel<<<blocks_for(m->rows), SETUP_T, 0, s>>>(m->rows, m->start, m->positions, diag_dev, info);
        e = cudaMemcpyAsync(h, info, sizeof(h), cudaMemcpyDeviceToHost, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(info);
    SMM_CUDA(e);
    *valid = h[0] == 0;
    *width = h[1];
    return SMM_OK;
}

bool smm_sgs_tiles_build_dev(smm_precond* p, const smm_csr* m, const int32_t* diag, int width) {
    const int rows = m->rows;
    if (rows < 2 * TILE || width > TILE_MAX_W) return false;
    if (width == 0) width = 1;
    cudaStream_t s = smm_default_stream();
    DevMem mem;
    int* flags = mem.get<int>(8);                              // [0] does not verify, [1] relaxation changed something, [2] highest level, [4..7] offsets
    if (!flags) return false;
    int hflags[8];
    auto read_flags = [&]() { return cudaMemcpyAsync(hflags, flags, sizeof(hflags), cudaMemcpyDeviceToHost, s) == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess; };
    if (cudaMemsetAsync(flags, 0, 8 * sizeof(int), s) != cudaSuccess) return false;
    // 1. proposal
    const int sample = 1 << 16;
    setup_offsets_kernel<<<blocks_for(3ll * sample), SETUP_T, 0, s>>>(rows, m->start, m->positions, sample, flags + 4);
    if (!read_flags() || hflags[7]) return false;
    std::vector<long long> offs;
    for (int i = 4; i < 7; ++i) if (hflags[i]) offs.push_back(hflags[i]);
    long long nx = 0, ny = 0, nz = 1;
    if (!smm_sgs_grid_from_offsets(rows, offs, &nx, &ny, &nz)) return false;
    int ti, tj, tk;
    smm_sgs_tile_shape(nz, &ti, &tj, &tk);
    const long long TI = (nx + ti - 1) / ti, TJ = (ny + tj - 1) / tj, TK = (nz + tk - 1) / tk;
    if (TI * TJ * TK >= (1ll << 25)) return false;             // positions are int32: 64 * tiles < 2^31
    const int ncl = (int)(TI * TJ * TK);
    const long long ntl = ncl, npos = ntl * TILE;
    // 2. rows of every tile
    int32_t* cl = mem.get<int32_t>((size_t)rows);
    int32_t* crow = mem.get<int32_t>((size_t)rows);
    int* cnt = mem.get<int>((size_t)ncl + 1);
    int* cptr = mem.get<int>((size_t)ncl + 1);
    int* cur = mem.get<int>((size_t)ncl + 1);
    int32_t* pred = mem.get<int32_t>((size_t)ncl * MAX_PREDS);
    uint8_t* npred = mem.get<uint8_t>((size_t)ncl);
    int* level = mem.get<int>((size_t)ncl);
    int32_t* tile_of = mem.get<int32_t>((size_t)ncl);
    int8_t* ilev = mem.get<int8_t>((size_t)rows);
    int32_t* where[2] = {mem.get<int32_t>((size_t)rows), mem.get<int32_t>((size_t)rows)};
    if (!cl || !crow || !cnt || !cptr || !cur || !pred || !npred || !level || !tile_of || !ilev || !where[0] || !where[1]) return false;
    if (cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)ncl + 1), s) != cudaSuccess) return false;
    setup_tile_of_row_kernel<<<blocks_for(rows), SETUP_T, 0, s>>>(rows, nx, ny, ti, tj, tk, TI, TJ, cl, cnt);
    if (!exclusive_scan(cnt, cptr, (long long)ncl + 1, mem, s)) return false;
    if (cudaMemcpyAsync(cur, cptr, sizeof(int) * ((size_t)ncl + 1), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return false;
    setup_scatter_kernel<<<blocks_for(rows), SETUP_T, 0, s>>>(rows, cl, cur, crow);
    setup_sort_rows_kernel<<<blocks_for(32ll * ncl), SETUP_T, 0, s>>>(ncl, cptr, crow, flags);
    // the arrays the sweep kernels keep
    p->threads_fwd = p->threads_bwd = npos;
    bool ok = cudaMalloc(&p->order_fwd, sizeof(int32_t) * npos) == cudaSuccess && cudaMalloc(&p->order_bwd, sizeof(int32_t) * npos) == cudaSuccess &&
              cudaMalloc(&p->ypos, sizeof(int32_t) * npos) == cudaSuccess && cudaMalloc(&p->yperm, sizeof(float) * npos) == cudaSuccess &&
              cudaMalloc(&p->xperm, sizeof(float) * npos) == cudaSuccess;
    const size_t esize = (size_t)ntl * width * TILE;
    for (int w = 0; w < 2 && ok; ++w) {
        p->esize[w] = (long long)esize;
        ok = cudaMalloc(&p->ecol[w], sizeof(int32_t) * esize) == cudaSuccess && cudaMalloc(&p->eidx[w], sizeof(int32_t) * esize) == cudaSuccess &&
             cudaMalloc(&p->tile_steps[w], (size_t)npos + (size_t)ntl) == cudaSuccess && cudaMalloc(&p->tile_push[w], sizeof(uint32_t) * npos) == cudaSuccess &&
             cudaMalloc(&p->eval[w], sizeof(float) * esize) == cudaSuccess && cudaMalloc(&p->dval[w], sizeof(float) * npos) == cudaSuccess;
    }
    int levels[2] = {0, 0};
    for (int w = 0; w < 2 && ok; ++w) {
        const bool forward = w == 0;
        const Sweep S{forward, m->start, m->positions, diag};
        int32_t* order = forward ? p->order_fwd : p->order_bwd;
        // 3. tile graph: predecessors, levels, order
        setup_preds_kernel<<<blocks_for(ncl), SETUP_T, 0, s>>>(S, ncl, cptr, crow, cl, pred, npred, flags);
        ok = cudaMemsetAsync(level, 0, sizeof(int) * (size_t)ncl, s) == cudaSuccess;
        const int max_rounds = 4096;                           // x 16 launches: a DAG of up to 65536 tile levels settles; a cyclic graph never does
        for (int round = 0; ok; ++round) {
            ok = cudaMemsetAsync(flags + 1, 0, 2 * sizeof(int), s) == cudaSuccess;
            for (int i = 0; i < 16; ++i) setup_relax_kernel<<<blocks_for(ncl), SETUP_T, 0, s>>>(ncl, pred, npred, level, flags + 1);
            setup_max_kernel<<<blocks_for(ncl), SETUP_T, 0, s>>>(ncl, level, flags + 2);
            ok = ok && read_flags() && hflags[0] == 0 && hflags[2] < 32768 && round < max_rounds;
            if (!hflags[1]) break;
        }
        if (!ok) break;
        const int nlev = hflags[2] + 1;
        levels[w] = nlev;
        const int nb = (ncl + HIST_BLOCK - 1) / HIST_BLOCK;
        if ((long long)nb * nlev > (64ll << 20)) { ok = false; break; }
        DevMem lmem;
        int* H = lmem.get<int>((size_t)nb * nlev);
        int* total = lmem.get<int>((size_t)nlev);
        int* lptr = lmem.get<int>((size_t)nlev);
        ok = H && total && lptr && cudaMemsetAsync(H, 0, sizeof(int) * (size_t)nb * nlev, s) == cudaSuccess;
        if (!ok) break;
        setup_level_hist_kernel<<<blocks_for(ncl), SETUP_T, 0, s>>>(ncl, forward, level, nlev, H);
        setup_level_colscan_kernel<<<blocks_for(nlev), SETUP_T, 0, s>>>(nb, nlev, H, total);
        setup_exclusive_small_kernel<<<1, 32, 0, s>>>(nlev, total, lptr);
        setup_tile_order_kernel<<<(unsigned)nb, HIST_BLOCK, 0, s>>>(ncl, forward, level, nlev, H, lptr, tile_of);
        if (forward) {                                         // host copy of the level boundaries (the device factorisations walk them)
            p->tile_level_ptr.assign((size_t)nlev + 1, ncl);
            ok = cudaMemcpyAsync(p->tile_level_ptr.data(), lptr, sizeof(int) * (size_t)nlev, cudaMemcpyDeviceToHost, s) == cudaSuccess;
            if (!ok) break;
        }
        // 4. inside the tiles
        ok = cudaMemsetAsync(order, 0xFF, sizeof(int32_t) * npos, s) == cudaSuccess &&                       // -1: padding
             cudaMemsetAsync(p->tile_steps[w], 0xFF, (size_t)npos + (size_t)ntl, s) == cudaSuccess &&        // 255: padding
             cudaMemsetAsync(where[w], 0, sizeof(int32_t) * (size_t)rows, s) == cudaSuccess &&
             cudaMemsetAsync(ilev, 0, (size_t)rows, s) == cudaSuccess &&
             cudaMemsetAsync(p->ecol[w], 0xFF, sizeof(int32_t) * esize, s) == cudaSuccess &&
             cudaMemsetAsync(p->eidx[w], 0xFF, sizeof(int32_t) * esize, s) == cudaSuccess;
        if (!ok) break;
        setup_steps_kernel<<<blocks_for(ncl, 64), 64, 0, s>>>(S, ncl, ntl, cptr, crow, cl, tile_of, ilev, order, p->tile_steps[w], where[w], flags);
        setup_entries_kernel<<<blocks_for(ncl, 64), 64, 0, s>>>(S, ncl, width, cptr, crow, tile_of, where[w], p->ecol[w], p->eidx[w], p->tile_push[w], flags);
        ok = cudaStreamSynchronize(s) == cudaSuccess;          // lmem goes away here
    }
    if (ok) {
        setup_ypos_kernel<<<blocks_for(npos), SETUP_T, 0, s>>>(npos, p->order_bwd, where[0], p->ypos);
        ok = read_flags() && hflags[0] == 0 && cudaGetLastError() == cudaSuccess;
    }
    if (!ok) { release_tiles(p); return false; }
    p->tile_blocks = 0;
    p->tile_width = width;
    p->tile_levels[0] = levels[0];
    p->tile_levels[1] = levels[1];
    p->tile_chain[0] = p->tile_chain[1] = 1;
    p->tiled = true;
    return true;
}
