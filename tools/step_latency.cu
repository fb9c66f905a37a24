// step_latency.cu -- what does one in-tile step of sgs_tiles.cu cost?  One warp: 128-bit shared load -> 4 products ->
// 4 dependent additions -> IEEE division -> 4 scattered shared stores -> __syncwarp, repeated; variants drop one piece.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 20000;
template <int MODE>
__global__ void steps(float* out, long long* cycles, float d, int clones) {
    __shared__ __align__(16) float stage[32][256];
    float* mine = stage[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    for (int i = lane; i < 256; i += 32) mine[i] = 1.0f + i * 1e-3f;
    __syncwarp();
    float v0 = 0.1f, v1 = 0.2f, v2 = 0.3f, v3 = 0.0f, init = 5.0f, keep = 0.f;
    const int p0 = (lane * 4 + 4) & 255, p1 = (lane * 4 + 33) & 255, p2 = (lane * 4 + 66) & 255, p3 = lane * 4;
    const long long t0 = clock64();
    for (int s = 0; s < N; ++s) {
        const float4 xo = *reinterpret_cast<const float4*>(mine + 4 * lane);
        float acc = init;
        acc = __fsub_rn(acc, __fmul_rn(v0, xo.x)); acc = __fsub_rn(acc, __fmul_rn(v1, xo.y));
        acc = __fsub_rn(acc, __fmul_rn(v2, xo.z)); acc = __fsub_rn(acc, __fmul_rn(v3, xo.w));
        float res = MODE == 1 ? acc * d : __fdiv_rn(acc, d);
        if ((s & 7) == (lane & 7) || MODE == 4) {
            if (MODE != 3) { mine[p0] = res; mine[p1] = res; mine[p2] = res; mine[p3] = res; }
            keep = res;
        }
        if (MODE != 2) __syncwarp();
    }
    const long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = keep;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 8);
    const char* names[5] = {"full step", "multiply instead of divide", "no __syncwarp", "no pushes", "all lanes store"};
    for (int warps : {1, 4, 8, 16, 32}) for (int m = 0; m < 5; ++m) {
        *cyc = 0;
        if (m == 0) steps<0><<<1, 32 * warps>>>(out, cyc, 6.0f, 0);
        if (m == 1) steps<1><<<1, 32 * warps>>>(out, cyc, 6.0f, 0);
        if (m == 2) steps<2><<<1, 32 * warps>>>(out, cyc, 6.0f, 0);
        if (m == 3) steps<3><<<1, 32 * warps>>>(out, cyc, 6.0f, 0);
        if (m == 4) steps<4><<<1, 32 * warps>>>(out, cyc, 6.0f, 0);
        cudaDeviceSynchronize();
        printf("%d warp(s)  %-28s %7.1f cycles per step\n", warps, names[m], (double)*cyc / N);
    }
    return 0;
}
