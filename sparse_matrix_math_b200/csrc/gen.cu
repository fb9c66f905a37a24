// gen.cu -- benchmark inputs generated directly in HBM (additive API; the reference can only ingest matrices
// through its std::map based TripletMatrix, which cannot hold the 1e8..1e9-entry configurations).
// Bit-identical to tests/matgen.py (tests/test_gpu_gen.py compares the arrays).
#include "smm_internal.cuh"

namespace {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---- stencils: row = (k*ny + j)*nx + i, columns ascending: -z -y -x diag +x +y +z --------------------------
struct StencilDesc {
    int nx, ny, nz;
    int use_y, use_z;
    float lo, diag, hi;
};

__device__ __forceinline__ int stencil_row_len(const StencilDesc& d, int i, int j, int k) {
    int len = 1 + (i > 0) + (i < d.nx - 1);
    if (d.use_y) len += (j > 0) + (j < d.ny - 1);
    if (d.use_z) len += (k > 0) + (k < d.nz - 1);
    return len;
}

// closed-form prefix: entries before row r = full*r minus the neighbours missing at the faces of rows < r
__device__ __forceinline__ long long stencil_start(const StencilDesc& d, long long r) {
    const long long nx = d.nx, ny = d.ny;
    const long long plane = nx * ny;
    const long long k = r / plane, rem = r % plane;
    const long long lines = r / nx, i = r % nx;
    const int full = 3 + 2 * d.use_y + 2 * d.use_z;
    long long miss = lines + (i > 0 ? 1 : 0);                 // rows < r with i == 0      (no -x neighbour)
    miss += lines;                                            // rows < r with i == nx-1   (no +x neighbour)
    if (d.use_y) {
        miss += k * nx + (rem < nx ? rem : nx);               // j == 0
        const long long last = rem - (ny - 1) * nx;
        miss += k * nx + (last > 0 ? last : 0);               // j == ny-1
    }
    if (d.use_z) {
        miss += (r < plane ? r : plane);                      // k == 0
        const long long last = r - (long long)(d.nz - 1) * plane;
        miss += (last > 0 ? last : 0);                        // k == nz-1
    }
    return r * full - miss;
}

// rows [row0, row0 + rows) of the global operator; start[] is relative to the first generated row
__global__ void stencil_kernel(const StencilDesc d, long long row0, long long rows, int32_t* __restrict__ start,
                               int32_t* __restrict__ positions, float* __restrict__ values) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > rows) return;
    const long long r = row0 + t;
    const long long s = stencil_start(d, r) - stencil_start(d, row0);
    start[t] = (int32_t)s;
    if (t == rows) return;
    const long long nx = d.nx, ny = d.ny, plane = nx * ny;
    const int k = (int)(r / plane), j = (int)((r % plane) / nx), i = (int)(r % nx);
    long long o = s;
    if (d.use_z && k > 0) { positions[o] = (int32_t)(r - plane); values[o] = d.lo; ++o; }
    if (d.use_y && j > 0) { positions[o] = (int32_t)(r - nx); values[o] = d.lo; ++o; }
    if (i > 0) { positions[o] = (int32_t)(r - 1); values[o] = d.lo; ++o; }
    positions[o] = (int32_t)r; values[o] = d.diag; ++o;
    if (i < d.nx - 1) { positions[o] = (int32_t)(r + 1); values[o] = d.hi; ++o; }
    if (d.use_y && j < d.ny - 1) { positions[o] = (int32_t)(r + nx); values[o] = d.hi; ++o; }
    if (d.use_z && k < d.nz - 1) { positions[o] = (int32_t)(r + plane); values[o] = d.hi; ++o; }
}

// ---- power law ---------------------------------------------------------------------------------------------
constexpr long long POWERLAW_A = 49152;

__device__ __forceinline__ int powerlaw_len(uint64_t seed, long long i, long long n, int cap) {
    const long long m = (long long)(splitmix64(seed, (uint64_t)i) >> 40) + 1;
    const long long q = (POWERLAW_A * POWERLAW_A) / m;
    long long l = (long long)floor(sqrt((double)q));
    if (l * l > q) --l;
    if ((l + 1) * (l + 1) <= q) ++l;                          // exact integer square root
    if (l > cap) l = cap;
    if (l > n) l = n;
    return (int)l;
}

__global__ void powerlaw_len_kernel(uint64_t seed, long long n, int cap, int32_t* __restrict__ len) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) len[i] = powerlaw_len(seed, i, n, cap);
}

// exclusive scan of n ints -> start[0..n] (three-pass: per-block sums, scan of block sums by one CTA, add)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void scan_tile_sums(const int32_t* __restrict__ in, long long n, long long* __restrict__ tile_sums) {
    __shared__ long long sh[SCAN_THREADS / 32];
    const long long base = (long long)blockIdx.x * SCAN_TILE;
    long long s = 0;
    for (int t = threadIdx.x; t < SCAN_TILE; t += SCAN_THREADS) {
        const long long i = base + t;
        if (i < n) s += in[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += sh[w];
        tile_sums[blockIdx.x] = t;
    }
}

__global__ void scan_tile_offsets(long long* tile_sums, long long ntiles) {
    // single thread: ntiles is n/4096 (2048 for 8M rows)
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long acc = 0;
        for (long long t = 0; t < ntiles; ++t) { const long long v = tile_sums[t]; tile_sums[t] = acc; acc += v; }
        tile_sums[ntiles] = acc;
    }
}

__global__ void scan_apply(const int32_t* __restrict__ in, long long n, const long long* __restrict__ tile_offs, int32_t* __restrict__ start) {
    // each thread owns SCAN_ITEMS consecutive items
    __shared__ long long sh[SCAN_THREADS];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) { const long long i = base + t; v[t] = i < n ? in[i] : 0; s += v[t]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long acc = tile_offs[blockIdx.x];
        for (int t = 0; t < SCAN_THREADS; ++t) { const long long x = sh[t]; sh[t] = acc; acc += x; }
    }
    __syncthreads();
    long long acc = sh[threadIdx.x];
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        const long long i = base + t;
        if (i < n) start[i] = (int32_t)acc;
        acc += v[t];
        if (i == n - 1) start[n] = (int32_t)acc;
    }
}

// one warp per row: lane-strided over the row's off-diagonal slots
__global__ void powerlaw_fill_kernel(uint64_t seed, long long n, const int32_t* __restrict__ start, int32_t* __restrict__ positions,
                                     float* __restrict__ values) {
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const long long s = start[row];
    const int l = start[row + 1] - start[row];
    const long long d = l - 1;
    const long long S = n - 1;
    const float lf = (float)l;
    int nbelow = 0;
    for (long long k = lane; k < d; k += 32) {
        const long long lo = (k * S) / d, hi = ((k + 1) * S) / d;
        const uint64_t key = ((uint64_t)row << 20) | (uint64_t)k;
        const uint64_t h1 = splitmix64(seed ^ 0xA5A5A5A5ull, key);
        const long long slot = lo + (long long)(h1 % (uint64_t)(hi - lo));
        const bool below = slot < row;
        const long long col = below ? slot : slot + 1;
        const uint64_t h2 = splitmix64(seed ^ 0x5A5A5A5Aull, key);
        const float u = __fsub_rn(__fdiv_rn((float)(h2 >> 40), 8388608.0f), 1.0f);
        const long long pos = s + k + (below ? 0 : 1);
        positions[pos] = (int32_t)col;
        values[pos] = __fdiv_rn(u, lf);
        nbelow += below ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) nbelow += __shfl_xor_sync(0xffffffffu, nbelow, o);
    if (lane == 0) {
        positions[s + nbelow] = (int32_t)row;
        values[s + nbelow] = 2.0f;
    }
}

__global__ void xstar_kernel(uint64_t seed, long long n, long long offset, float* __restrict__ x) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = __fdiv_rn((float)(splitmix64(seed, (uint64_t)(i + offset)) >> 40), 16777216.0f);
}

}  // namespace

extern "C" {

static int gen_stencil_rows(int kind, int nx, int ny, int nz, float c, long long row0, long long row1, smm_csr_t** out) {
    cudaStream_t s = smm_default_stream();
    StencilDesc d;
    d.nx = nx; d.ny = ny; d.nz = kind == SMM_GEN_POISSON2D ? 1 : nz;
    if (d.nx < 1 || d.ny < 1 || d.nz < 1) { smm_set_error("smm_gen_csr: bad grid"); return SMM_E_INVALID; }
    d.use_y = 1; d.use_z = kind == SMM_GEN_CONVDIFF3D;
    if (kind == SMM_GEN_POISSON2D) { d.lo = -1.0f; d.hi = -1.0f; d.diag = 4.0f; }
    else { d.lo = -1.0f - c; d.hi = -1.0f + c; d.diag = 6.0f; }
    const long long total = (long long)d.nx * d.ny * d.nz;
    if (row0 < 0 || row1 < row0 || row1 > total) { smm_set_error("smm_gen_csr_rows: bad row range"); return SMM_E_INVALID; }
    const long long rows = row1 - row0;
    if (total > 0x7fffffffll - 1) { smm_set_error("smm_gen_csr: exceeds 32-bit indices"); return SMM_E_INVALID; }
    int32_t* start = nullptr;
    SMM_CUDA(cudaMalloc(&start, sizeof(int32_t) * (size_t)(rows + 1)));
    // two passes: row offsets first (to learn nnz), then the entries
    const int full = 3 + 2 * d.use_y + 2 * d.use_z;
    const size_t nmax = (((size_t)rows * full) + 3) & ~(size_t)3;
    int32_t* positions = nullptr;
    float* values = nullptr;
    SMM_CUDA(cudaMalloc(&positions, sizeof(int32_t) * (nmax ? nmax : 4)));
    SMM_CUDA(cudaMalloc(&values, sizeof(float) * (nmax ? nmax : 4)));
    stencil_kernel<<<(unsigned)((rows + 1 + 255) / 256), 256, 0, s>>>(d, row0, rows, start, positions, values);
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    SMM_CUDA(cudaStreamSynchronize(s));
    int rc = smm_csr_create_dev((int)rows, (int)total, start, positions, values, 0, out);
    if (rc == SMM_OK) (*out)->nnz_alloc = (int64_t)(nmax ? nmax : 4);       // our own padded allocation
    return rc;
}

int smm_gen_csr_rows(int kind, int nx, int ny, int nz, float c, int64_t row_begin, int64_t row_end, smm_csr_t** out) {
    if (!out || (kind != SMM_GEN_POISSON2D && kind != SMM_GEN_CONVDIFF3D)) return SMM_E_INVALID;
    return gen_stencil_rows(kind, nx, ny, nz, c, row_begin, row_end, out);
}

int smm_gen_csr(int kind, int nx, int ny, int nz, float c, uint64_t seed, smm_csr_t** out) {
    if (!out) return SMM_E_INVALID;
    cudaStream_t s = smm_default_stream();
    int32_t *start = nullptr, *positions = nullptr;
    float* values = nullptr;
    long long rows = 0, nnz = 0;
    if (kind == SMM_GEN_POISSON2D || kind == SMM_GEN_CONVDIFF3D) {
        const long long total = (long long)nx * ny * (kind == SMM_GEN_POISSON2D ? 1 : nz);
        return gen_stencil_rows(kind, nx, ny, nz, c, 0, total, out);
    } else if (kind == SMM_GEN_POWERLAW) {
        rows = nx;
        const int cap = ny > 1 ? ny : 131072;          // ny = optional row-length cap
        if (rows < 2 || rows >= (1ll << 31) - 1) { smm_set_error("smm_gen_csr: bad powerlaw size"); return SMM_E_INVALID; }
        int32_t* len = nullptr;
        long long* tiles = nullptr;
        const long long ntiles = (rows + SCAN_TILE - 1) / SCAN_TILE;
        SMM_CUDA(cudaMalloc(&len, sizeof(int32_t) * (size_t)rows));
        SMM_CUDA(cudaMalloc(&tiles, sizeof(long long) * (size_t)(ntiles + 1)));
        SMM_CUDA(cudaMalloc(&start, sizeof(int32_t) * (size_t)(rows + 1)));
        powerlaw_len_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(seed, rows, cap, len);
        scan_tile_sums<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(len, rows, tiles);
        scan_tile_offsets<<<1, 32, 0, s>>>(tiles, ntiles);
        scan_apply<<<(unsigned)ntiles, SCAN_THREADS, 0, s>>>(len, rows, tiles, start);
        SMM_COUNT_LAUNCH(4);
        SMM_CUDA(cudaGetLastError());
        SMM_CUDA(cudaMemcpyAsync(&nnz, tiles + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, s));
        SMM_CUDA(cudaStreamSynchronize(s));
        cudaFree(len);
        cudaFree(tiles);
        if (nnz > 0x7fffffffll) { cudaFree(start); smm_set_error("smm_gen_csr: exceeds 32-bit indices"); return SMM_E_INVALID; }
        const size_t npad = ((size_t)nnz + 3) & ~(size_t)3;
        SMM_CUDA(cudaMalloc(&positions, sizeof(int32_t) * npad));
        SMM_CUDA(cudaMalloc(&values, sizeof(float) * npad));
        powerlaw_fill_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, s>>>(seed, rows, start, positions, values);
        SMM_COUNT_LAUNCH(1);
        SMM_CUDA(cudaGetLastError());
    } else {
        smm_set_error("smm_gen_csr: unknown kind %d", kind);
        return SMM_E_INVALID;
    }
    SMM_CUDA(cudaStreamSynchronize(s));
    {
        int rc = smm_csr_create_dev((int)rows, (int)rows, start, positions, values, 0, out);
        if (rc == SMM_OK) (*out)->nnz_alloc = (int64_t)((((size_t)nnz + 3) & ~(size_t)3));
        return rc;
    }
}

int smm_gen_xstar_dev(int64_t n, int64_t offset, uint64_t seed, float* x_dev, void* stream) {
    if (n < 0 || (n && !x_dev)) return SMM_E_INVALID;
    if (n == 0) return SMM_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : smm_default_stream();
    xstar_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(seed, n, offset, x_dev);
    SMM_COUNT_LAUNCH(1);
    SMM_CUDA(cudaGetLastError());
    return SMM_OK;
}

}  // extern "C"
