// dots.cu -- stand-alone dot products in the reference's summation orders (parity modes), plus the C-ABI smm_dot.
//
//   SMM_REDUCE_REFERENCE_TREE   Vector::operator* of the SMM_MULTITHREADING build (H:308-320):
//       tbb::parallel_deterministic_reduce over blocked_range<int>(0,n,8192), identity 0.0f, std::plus.
//       The range is halved at begin+(end-begin)/2 while its size exceeds the grain; each leaf is summed left to
//       right from 0; joins are left+right.  All nodes at depth D' = min{d : floor(n/2^d) <= 8192} exist, and each
//       is either a leaf or (size 8193) splits exactly once more, so: one WARP per depth-D' node walks down from
//       the root to find its range, sums its one or two leaves (coalesced loads by all lanes, the sequential adds by
//       lane 0 out of shared memory), and the 2^D' node values are then combined by a perfect pairwise tree -- the
//       same additions in the same order as the reference.
//   SMM_REDUCE_REFERENCE_SERIAL the serial build (H:322-326): left to right.  One thread adds, the rest of its CTA
//       streams products into shared memory ahead of it.
//   SMM_REDUCE_FAST             vecops.cu's fused two-stage reduction (VEC_DOT2).
#include "epilogue.cuh"
#include "smm_internal.cuh"

namespace {

constexpr int TBB_GRAIN = 8192;

constexpr int LEAF_CHUNK = 512;        // products staged per round: 16 per lane
constexpr int TREE_WARPS = 4;

// One warp sums one leaf: all lanes stream a[], b[] with coalesced loads (the next chunk is already in flight in
// registers), multiply, and park the products in shared memory; lane 0 then adds them left to right from 0, which
// is the order of the reference's leaf loop (H:312-316).  The 4-cycle FADD chain of lane 0 (8192 adds per leaf) is
// the critical path; the loads hide under it.
__device__ __forceinline__ float leaf_sum(const float* __restrict__ a, const float* __restrict__ b, long long lo, long long hi, float* buf) {
    const int lane = threadIdx.x & 31;
    float cur = 0.0f;                                        // identity, H:312
    float ra[LEAF_CHUNK / 32], rb[LEAF_CHUNK / 32];
    auto fetch = [&](long long base) {
#pragma unroll
        for (int k = 0; k < LEAF_CHUNK / 32; ++k) {
            const long long j = base + k * 32 + lane;
            const bool in = j < hi;
            ra[k] = in ? __ldg(a + j) : 0.0f;
            rb[k] = in ? __ldg(b + j) : 0.0f;
        }
    };
    fetch(lo);
    for (long long base = lo; base < hi; base += LEAF_CHUNK) {
#pragma unroll
        for (int k = 0; k < LEAF_CHUNK / 32; ++k) buf[k * 32 + lane] = __fmul_rn(ra[k], rb[k]);
        __syncwarp();
        if (base + LEAF_CHUNK < hi) fetch(base + LEAF_CHUNK);
        if (lane == 0) {
            const long long left = hi - base;
            const int m = left < LEAF_CHUNK ? (int)left : LEAF_CHUNK;
            int t = 0;
            for (; t + 4 <= m; t += 4) {
                const float4 q = *reinterpret_cast<const float4*>(buf + t);
                cur = __fadd_rn(cur, q.x); cur = __fadd_rn(cur, q.y); cur = __fadd_rn(cur, q.z); cur = __fadd_rn(cur, q.w);   // H:314-316
            }
            for (; t < m; ++t) cur = __fadd_rn(cur, buf[t]);
        }
        __syncwarp();
    }
    return cur;                                              // valid in lane 0
}

struct TreeParams {
    long long n;
    int depth;             // D'
    int ndots;             // 1 or 2
    const float* a[2];
    const float* b[2];
    float* nodes;          // [2 dots][2 ping-pong][2^D']
    unsigned int* ticket;
    SolveState* state;
    int finish;
    float* out_dev;        // optional: totals written here too
};

__global__ void __launch_bounds__(TREE_WARPS * 32) dot_tree_kernel(const TreeParams P) {
    if (P.state != nullptr && P.state->done) return;
    __shared__ int sh_last;
    __shared__ __align__(16) float sh_buf[TREE_WARPS][LEAF_CHUNK];
    const long long nn = 1ll << P.depth;
    const int warp = threadIdx.x >> 5;
    const long long job = (long long)blockIdx.x * TREE_WARPS + warp;     // (dot, depth-D' node)
    if (job < nn * P.ndots) {
        const int d = (int)(job / nn);
        const long long i = job - (long long)d * nn;
        long long lo = 0, hi = P.n;
        for (int level = P.depth - 1; level >= 0; --level) {
            const long long mid = lo + (hi - lo) / 2;
            if ((i >> level) & 1) lo = mid; else hi = mid;
        }
        float v;
        if (hi - lo > TBB_GRAIN) {
            const long long mid = lo + (hi - lo) / 2;
            const float l = leaf_sum(P.a[d], P.b[d], lo, mid, sh_buf[warp]);
            v = __fadd_rn(l, leaf_sum(P.a[d], P.b[d], mid, hi, sh_buf[warp]));
        } else {
            v = leaf_sum(P.a[d], P.b[d], lo, hi, sh_buf[warp]);
        }
        if ((threadIdx.x & 31) == 0) P.nodes[(size_t)d * 2 * nn + i] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) sh_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    // perfect pairwise tree, ping-pong between the two halves of each dot's node buffer
    float totals[2] = {0.f, 0.f};
    for (int d = 0; d < P.ndots; ++d) {
        float* cur = P.nodes + (size_t)d * 2 * nn;
        float* nxt = cur + nn;
        for (long long w = nn >> 1; w >= 1; w >>= 1) {
            for (long long k = threadIdx.x; k < w; k += blockDim.x) {
                const float2 pr = __ldcg(reinterpret_cast<const float2*>(cur) + k);
                nxt[k] = __fadd_rn(pr.x, pr.y);
            }
            __threadfence_block();
            __syncthreads();
            float* t = cur; cur = nxt; nxt = t;
        }
        totals[d] = __ldcg(cur);
    }
    if (threadIdx.x == 0) {
        *P.ticket = 0u;
        if (P.out_dev) { P.out_dev[0] = totals[0]; P.out_dev[1] = totals[1]; }
        if (P.state) smm_finish(P.finish, P.state, totals[0], totals[1]);
    }
}

struct SerialParams {
    long long n;
    int ndots;
    const float* a[2];
    const float* b[2];
    SolveState* state;
    int finish;
    float* out_dev;
};

constexpr int SER_THREADS = 256;
constexpr int SER_CHUNK = 4096;

__global__ void __launch_bounds__(SER_THREADS) dot_serial_kernel(const SerialParams P) {
    if (P.state != nullptr && P.state->done) return;
    __shared__ float buf[2][SER_CHUNK];
    float totals[2] = {0.f, 0.f};
    for (int d = 0; d < P.ndots; ++d) {
        const float* a = P.a[d];
        const float* b = P.b[d];
        float cur = 0.0f;
        const long long nchunks = (P.n + SER_CHUNK - 1) / SER_CHUNK;
        // prologue: chunk 0
        for (int t = threadIdx.x; t < SER_CHUNK; t += SER_THREADS) {
            const long long j = t;
            buf[0][t] = j < P.n ? __fmul_rn(a[j], b[j]) : 0.0f;
        }
        __syncthreads();
        for (long long c = 0; c < nchunks; ++c) {
            const int pb = (int)(c & 1);
            if (threadIdx.x >= 32) {                          // warps 1..7 prefetch the next chunk
                if (c + 1 < nchunks) {
                    for (int t = threadIdx.x - 32; t < SER_CHUNK; t += SER_THREADS - 32) {
                        const long long j = (c + 1) * SER_CHUNK + t;
                        buf[pb ^ 1][t] = j < P.n ? __fmul_rn(a[j], b[j]) : 0.0f;
                    }
                }
            } else if (threadIdx.x == 0) {                    // one thread adds left to right, H:322-326
                const long long left = P.n - c * SER_CHUNK;
                const int m = left < SER_CHUNK ? (int)left : SER_CHUNK;
                for (int t = 0; t < m; ++t) cur = __fadd_rn(cur, buf[pb][t]);
            }
            __syncthreads();
        }
        totals[d] = cur;
    }
    if (threadIdx.x == 0) {
        if (P.out_dev) { P.out_dev[0] = totals[0]; P.out_dev[1] = totals[1]; }
        if (P.state) smm_finish(P.finish, P.state, totals[0], totals[1]);
    }
}

struct DotScratch {
    float* nodes = nullptr;
    size_t nodes_cap = 0;
    unsigned int* ticket = nullptr;
    float* out = nullptr;        // 2 floats
};
DotScratch g_scratch[64];
}  // namespace
int smm_tree_depth(long long n);
namespace {

int scratch_for(long long nn, DotScratch** out) {
    int dev = 0;
    SMM_CUDA(cudaGetDevice(&dev));
    DotScratch& sc = g_scratch[dev];
    if (!sc.ticket) {
        SMM_CUDA(cudaMalloc(&sc.ticket, sizeof(unsigned int)));
        SMM_CUDA(cudaMemset(sc.ticket, 0, sizeof(unsigned int)));
        SMM_CUDA(cudaMalloc(&sc.out, 2 * sizeof(float)));
    }
    const size_t need = (size_t)nn * 4;
    if (sc.nodes_cap < need) {
        cudaFree(sc.nodes);
        SMM_CUDA(cudaMalloc(&sc.nodes, need * sizeof(float)));
        sc.nodes_cap = need;
    }
    *out = &sc;
    return SMM_OK;
}

}  // namespace

// allocate the node scratch for vectors of length n now (cudaMalloc is not allowed while a stream is capturing)
int smm_dot_ref_prepare(long long n) {
    DotScratch* sc = nullptr;
    return scratch_for(1ll << smm_tree_depth(n), &sc);
}

int smm_tree_depth(long long n) {
    int d = 0;
    while ((n >> d) > TBB_GRAIN) ++d;                        // floor(n / 2^d) <= grain
    return d;
}

// t0 = a0.b0 [, t1 = a1.b1] in the requested reference order; the totals go to smm_finish(finish, state, t0, t1)
// and/or out_dev[0..1].  mode: SMM_REDUCE_REFERENCE_TREE or SMM_REDUCE_REFERENCE_SERIAL.
int smm_launch_dot_ref(int mode, long long n, int ndots, const float* a0, const float* b0, const float* a1, const float* b1,
                       SolveState* state, int finish, float* out_dev, cudaStream_t s) {
    if (mode == SMM_REDUCE_REFERENCE_TREE) {
        TreeParams P;
        P.n = n; P.depth = smm_tree_depth(n); P.ndots = ndots;
        P.a[0] = a0; P.b[0] = b0; P.a[1] = a1; P.b[1] = b1;
        DotScratch* sc = nullptr;
        SMM_TRY(scratch_for(1ll << P.depth, &sc));
        P.nodes = sc->nodes; P.ticket = sc->ticket; P.state = state; P.finish = finish; P.out_dev = out_dev;
        const long long jobs = (1ll << P.depth) * ndots;       // one warp per (dot, depth-D' node)
        dot_tree_kernel<<<(unsigned)((jobs + TREE_WARPS - 1) / TREE_WARPS), TREE_WARPS * 32, 0, s>>>(P);
    } else {
        SerialParams P;
        P.n = n; P.ndots = ndots;
        P.a[0] = a0; P.b[0] = b0; P.a[1] = a1; P.b[1] = b1;
        P.state = state; P.finish = finish; P.out_dev = out_dev;
        dot_serial_kernel<<<1, SER_THREADS, 0, s>>>(P);
    }
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

// ---------------------------------------------------------------------------------------------------
// C ABI: Vector::operator* (H:305-328)
// ---------------------------------------------------------------------------------------------------
namespace {
smm_workspace* g_dot_ws[64] = {nullptr};

int dot_workspace(smm_workspace** out) {
    int dev = 0;
    SMM_CUDA(cudaGetDevice(&dev));
    if (!g_dot_ws[dev]) {
        smm_csr fake;                                          // matrix-free workspace
        fake.device = dev;
        fake.num_blocks = 0;
        smm_workspace* ws = nullptr;
        SMM_TRY(smm_workspace_get(&fake, &ws));
        g_dot_ws[dev] = ws;
        fake.ws = nullptr;
    }
    *out = g_dot_ws[dev];
    return SMM_OK;
}
}  // namespace

extern "C" {

int smm_dot_dev(int64_t n, const float* a_dev, const float* b_dev, int reduction_mode, float* out_host, void* stream) {
    if (n < 0 || !out_host || (n && (!a_dev || !b_dev))) { smm_set_error("smm_dot: bad arguments"); return SMM_E_INVALID; }
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    float res[2] = {0.f, 0.f};
    if (reduction_mode == SMM_REDUCE_FAST) {
        smm_workspace* ws = nullptr;
        SMM_TRY(dot_workspace(&ws));
        SMM_CUDA(cudaMemsetAsync(ws->state, 0, sizeof(SolveState), s));
        VecArgs v;
        v.n = n; v.in[0] = a_dev; v.in[1] = b_dev; v.state = ws->state; v.finish = FIN_STORE; v.slot = 0; v.ws = ws;
        SMM_TRY(smm_launch_vec(VEC_DOT2, v, s));
        SMM_CUDA(cudaMemcpyAsync(res, (const char*)ws->state + offsetof(SolveState, scratch), 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
    } else if (reduction_mode == SMM_REDUCE_REFERENCE_TREE || reduction_mode == SMM_REDUCE_REFERENCE_SERIAL) {
        DotScratch* sc = nullptr;
        SMM_TRY(scratch_for(1ll << smm_tree_depth(n), &sc));
        SMM_TRY(smm_launch_dot_ref(reduction_mode, n, 1, a_dev, b_dev, a_dev, b_dev, nullptr, FIN_NONE, sc->out, s));
        SMM_CUDA(cudaMemcpyAsync(res, sc->out, 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
    } else {
        smm_set_error("smm_dot: unknown reduction mode %d", reduction_mode);
        return SMM_E_INVALID;
    }
    *out_host = res[0];
    return SMM_OK;
}

int smm_dot(int64_t n, const float* a, const float* b, int reduction_mode, float* out) {
    if (n < 0 || !out || (n && (!a || !b))) { smm_set_error("smm_dot: bad arguments"); return SMM_E_INVALID; }
    float *da = nullptr, *db = nullptr;
    const size_t bytes = sizeof(float) * (size_t)(n ? n : 1);
    SMM_CUDA(cudaMalloc(&da, bytes));
    if (cudaMalloc(&db, bytes) != cudaSuccess) { cudaFree(da); return smm_cuda_fail(cudaGetLastError(), "cudaMalloc", __FILE__, __LINE__); }
    int rc = SMM_OK;
    if (n && (cudaMemcpy(da, a, bytes, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(db, b, bytes, cudaMemcpyHostToDevice) != cudaSuccess))
        rc = smm_cuda_fail(cudaGetLastError(), "memcpy", __FILE__, __LINE__);
    if (rc == SMM_OK) rc = smm_dot_dev(n, da, db, reduction_mode, out, nullptr);
    cudaFree(da);
    cudaFree(db);
    return rc;
}

}  // extern "C"
