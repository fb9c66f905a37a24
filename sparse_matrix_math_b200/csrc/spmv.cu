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
// Algorithmic bytes per launch: 8*nnz (values+positions) + 4*(rows+1) (start) + 4*cols (mult, gathered once)
// + 4*rows (out) [+ 4*rows lhs for ADD/SUB] [+ 4*rows per fused-dot operand].
#include "smm_internal.cuh"

namespace {

struct SpmvParams {
    const int32_t* __restrict__ start;
    const int32_t* __restrict__ positions;
    const float* __restrict__ values;
    const int32_t* __restrict__ block_row;
    int rows;
    int nnz;
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
};

}  // namespace

#include "epilogue.cuh"

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

__global__ void __launch_bounds__(SPMV_THREADS) spmv_kernel(const SpmvParams P) {
    if (P.state != nullptr && P.state->done) return;

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
            // two vectors per thread in flight before the dependent gathers
            for (int v = tid; v < nvec; v += 2 * SPMV_THREADS) {
                const int v2 = v + SPMV_THREADS;
                const bool has2 = v2 < nvec;
                const int4 c = ldg_stream_i4(pos4 + v);
                const float4 a = ldg_stream_f4(val4 + v);
                int4 c2 = make_int4(0, 0, 0, 0);
                float4 a2 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (has2) { c2 = ldg_stream_i4(pos4 + v2); a2 = ldg_stream_f4(val4 + v2); }
                const float x0 = __ldg(P.mult + c.x), x1 = __ldg(P.mult + c.y), x2 = __ldg(P.mult + c.z), x3 = __ldg(P.mult + c.w);
                float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
                if (has2) { y0 = __ldg(P.mult + c2.x); y1 = __ldg(P.mult + c2.y); y2 = __ldg(P.mult + c2.z); y3 = __ldg(P.mult + c2.w); }
                reinterpret_cast<float4*>(prod)[v] = make_float4(__fmul_rn(a.x, x0), __fmul_rn(a.y, x1), __fmul_rn(a.z, x2), __fmul_rn(a.w, x3));
                if (has2) reinterpret_cast<float4*>(prod)[v2] = make_float4(__fmul_rn(a2.x, y0), __fmul_rn(a2.y, y1), __fmul_rn(a2.z, y2), __fmul_rn(a2.w, y3));
            }
            {   // tail of the window (fewer than 4 entries)
                const int k = a0 + (nvec << 2) + tid;
                if (k < k1) prod[k - a0] = __fmul_rn(ldg_stream_f(P.values + k), __ldg(P.mult + ldg_stream_i(P.positions + k)));
            }
            __syncthreads();
            bool saw_long = false;
            for (int r = r0 + tid; r < r1; r += SPMV_THREADS) {
                const int s = P.start[r] - a0, e = P.start[r + 1] - a0;
                if (!P.exact && e - s > SPMV_LONG_IN_STREAM) { saw_long = true; continue; }
                float dot = 0.0f;                             // H:1484 ; empty row -> op(lhs, 0), H:1479-1483
                for (int j = s; j < e; ++j) dot = __fadd_rn(prod[j], dot);   // H:1485-1489, `val*x + dot`, two roundings
                write(r, dot);
            }
            if (saw_long) sh_long = 1;
            __syncthreads();
            if (sh_long) {
                for (int r = r0 + warp; r < r1; r += NWARPS) {
                    const int s = P.start[r] - a0, e = P.start[r + 1] - a0;
                    if (e - s <= SPMV_LONG_IN_STREAM) continue;
                    float acc = 0.0f;
                    for (int j = s + lane; j < e; j += 32) acc += prod[j];
                    acc = warp_sum(acc);
                    if (lane == 0) write(r, acc);
                }
            }
        } else if (P.exact) {
            // ---------------- exact row path: one thread per row, left to right from global memory ----------------
            for (int r = r0 + tid; r < r1; r += SPMV_THREADS) {
                const int s = P.start[r], e = P.start[r + 1];
                float dot = 0.0f;
                for (int k = s; k < e; ++k) dot = __fadd_rn(__fmul_rn(P.values[k], __ldg(P.mult + P.positions[k])), dot);
                write(r, dot);
            }
        } else {
            // ---------------- row path: warp per row, CTA per very long row ----------------
            bool any_cta_row = false;
            for (int r = r0 + warp; r < r1; r += NWARPS) {
                const int s = P.start[r], e = P.start[r + 1];
                if (e - s > SPMV_CTA_ROW) { any_cta_row = true; continue; }
                float acc = 0.0f;
                int k = s + lane;
                for (; k + 32 < e; k += 64) {                 // two independent gathers in flight
                    const int c0 = ldg_stream_i(P.positions + k), c1 = ldg_stream_i(P.positions + k + 32);
                    const float a0v = ldg_stream_f(P.values + k), a1v = ldg_stream_f(P.values + k + 32);
                    acc = fmaf(a0v, __ldg(P.mult + c0), acc);
                    acc = fmaf(a1v, __ldg(P.mult + c1), acc);
                }
                if (k < e) acc = fmaf(ldg_stream_f(P.values + k), __ldg(P.mult + ldg_stream_i(P.positions + k)), acc);
                acc = warp_sum(acc);
                if (lane == 0) write(r, acc);
            }
            // nrows < SPMV_MIN_STREAM_ROWS or a row longer than the window: rescan for CTA-wide rows (uniform branch)
            if (__syncthreads_or(any_cta_row)) {
                for (int r = r0; r < r1; ++r) {
                    const int s = P.start[r], e = P.start[r + 1];
                    if (e - s <= SPMV_CTA_ROW) continue;
                    float acc = 0.0f;
                    const int sa = (s + 3) & ~3;              // aligned body, scalar head/tail
                    const int ea = e & ~3;
                    if (tid < sa - s) acc = fmaf(P.values[s + tid], __ldg(P.mult + P.positions[s + tid]), acc);
                    if (tid < e - ea) acc = fmaf(P.values[ea + tid], __ldg(P.mult + P.positions[ea + tid]), acc);
                    const int4* pos4 = reinterpret_cast<const int4*>(P.positions + sa);
                    const float4* val4 = reinterpret_cast<const float4*>(P.values + sa);
                    const int nvec = (ea - sa) >> 2;
                    for (int v = tid; v < nvec; v += SPMV_THREADS) {
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
    }

    if (P.reduce != RED_NONE) {
        float v[2] = {write.acc0, write.acc1};
        __syncthreads();
        if (grid_sum_last_block<2>(v, P.partials, P.partials_stride, P.ticket, red_sh, &sh_flag)) {
            if (tid == 0) smm_finish(P.finish, P.state, v[0], v[1]);
        }
    }
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

__global__ void first_active_kernel(const int32_t* __restrict__ start, int rows, int* out) {
    // firstActiveStart (H:1622-1628): first row i with start[i+1] != 0, or rows
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows && start[i + 1] != 0 && start[i] == 0) atomicMin(out, i);
}

}  // namespace

int smm_csr_analyse(smm_csr* m, cudaStream_t s) {
    m->num_blocks = (int)(m->nnz / SPMV_CHUNK) + 1;
    if (m->block_row) { cudaFree(m->block_row); m->block_row = nullptr; }
    SMM_CUDA(cudaMalloc(&m->block_row, sizeof(int32_t) * (size_t)(m->num_blocks + 1)));
    const int threads = 256;
    block_row_kernel<<<(m->num_blocks + 1 + threads - 1) / threads, threads, 0, s>>>(m->start, m->rows, m->num_blocks, m->block_row);
    SMM_COUNT_LAUNCH(1);
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

int smm_launch_spmv(const SpmvArgs& a, cudaStream_t s) {
    const smm_csr* m = a.m;
    if (m->rows == 0) return SMM_OK;
    SpmvParams P;
    P.start = m->start; P.positions = m->positions; P.values = m->values; P.block_row = m->block_row;
    P.rows = m->rows; P.nnz = (int)m->nnz; P.op = a.op; P.exact = a.exact;
    P.lhs = a.lhs; P.mult = a.mult; P.out = a.out;
    P.copy1 = a.copy1; P.copy2 = a.copy2; P.copy3 = a.copy3;
    P.aux = a.aux; P.reduce = a.reduce; P.finish = a.finish; P.state = a.state;
    P.partials = nullptr; P.partials_stride = 0; P.ticket = nullptr;
    if (a.reduce != RED_NONE) {
        smm_workspace* ws = m->ws;
        if (!ws || ws->partials_cap < (size_t)m->num_blocks) { smm_set_error("spmv: reduction workspace too small"); return SMM_E_STATE; }
        P.partials = ws->partials + (size_t)a.slot * 2 * ws->partials_cap;
        P.partials_stride = ws->partials_cap;
        P.ticket = ws->tickets + a.slot;
    }
    spmv_kernel<<<m->num_blocks, SPMV_THREADS, 0, s>>>(P);
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}
