// gather_ceiling.cu -- how fast can a B200 do what the irregular-row SpMV's inner loop does: stream an index (and a value) per
// entry with coalesced 128-bit loads and gather x[index] from an L2-resident vector, indices without any locality?
// The rate measured here is the empirical ceiling quoted beside config 4's SpMV.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/bin/gather_ceiling tools/gather_ceiling.cu
//   tools/bin/gather_ceiling [vector elements = 8388608] [gathers = 197157160]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void fill_kernel(int4* idx, float4* val, long long n4, unsigned int cols) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x5EEDull;
        int c[4];
        for (int k = 0; k < 4; ++k) {
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
            c[k] = (int)(z % cols);
        }
        idx[i] = make_int4(c[0], c[1], c[2], c[3]);
        val[i] = make_float4(1.f, 0.5f, 0.25f, 0.125f);
    }
}

// WITH_VALUES: also stream a value per entry and multiply (the SpMV's 8 bytes per entry); DEPTH 128-bit vectors in flight per thread
template <bool WITH_VALUES, int DEPTH>
__global__ void __launch_bounds__(256) gather_kernel(const int4* __restrict__ idx, const float4* __restrict__ val, const float* __restrict__ x, long long n4, float* out) {
    float acc = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += DEPTH * stride) {
        int4 c[DEPTH];
        float4 a[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const long long j = i + d * stride;
            c[d] = j < n4 ? __ldcs(idx + j) : make_int4(0, 0, 0, 0);
            a[d] = (WITH_VALUES && j < n4) ? __ldcs(val + j) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
            acc += a[d].x * __ldg(x + c[d].x) + a[d].y * __ldg(x + c[d].y) + a[d].z * __ldg(x + c[d].z) + a[d].w * __ldg(x + c[d].w);
    }
    if (acc == 12345.678f) out[0] = acc;      // keep the loads
}

template <bool WV, int DEPTH>
float run(const int4* idx, const float4* val, const float* x, long long n4, float* out, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) gather_kernel<WV, DEPTH><<<grid, 256>>>(idx, val, x, n4, out);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int r = 0; r < reps; ++r) gather_kernel<WV, DEPTH><<<grid, 256>>>(idx, val, x, n4, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main(int argc, char** argv) {
    const unsigned int cols = argc > 1 ? (unsigned int)atoll(argv[1]) : 8388608u;
    const long long n = argc > 2 ? atoll(argv[2]) : 197157160ll;
    const long long n4 = n / 4;
    int4* idx; float4* val; float* x; float* out;
    cudaMalloc(&idx, sizeof(int4) * n4); cudaMalloc(&val, sizeof(float4) * n4); cudaMalloc(&x, sizeof(float) * cols); cudaMalloc(&out, 4);
    cudaMemset(x, 0, sizeof(float) * cols);
    fill_kernel<<<1184, 256>>>(idx, val, n4, cols);
    cudaDeviceSynchronize();
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    printf("%s: %lld gathers of 4 bytes from a vector of %u floats (%.1f MB), indices hashed (no locality)\n", prop.name, n4 * 4, cols, cols * 4e-6);
    for (int per_sm : {4, 8}) {
        const int grid = prop.multiProcessorCount * per_sm;
        const float a2 = run<false, 2>(idx, val, x, n4, out, grid), a4 = run<false, 4>(idx, val, x, n4, out, grid);
        const float b2 = run<true, 2>(idx, val, x, n4, out, grid), b4 = run<true, 4>(idx, val, x, n4, out, grid);
        printf("%d CTAs per SM: index + gather       %.3f / %.3f ms (2 / 4 vectors in flight) = %.0f Ggather/s\n", per_sm, a2, a4, n4 * 4 / (a2 < a4 ? a2 : a4) * 1e-6);
        printf("%d CTAs per SM: index + value + gather %.3f / %.3f ms                          = %.0f Ggather/s\n", per_sm, b2, b4, n4 * 4 / (b2 < b4 ? b2 : b4) * 1e-6);
    }
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
