// hop_latency.cu -- how long does one producer->consumer hand-off through memory take on this GPU?
// The sync-free triangular sweeps (csrc/sgs.cu) are bounded by (number of levels) x (this latency), so the floor is
// measured here for every mechanism that could carry the hand-off:
//   gmem     st.relaxed.gpu / ld.relaxed.gpu polling between two CTAs (what sgs.cu does)
//   gmem_vol volatile store / load
//   gmem_acq st.release.gpu / ld.acquire.gpu
//   atom     red.add publish / atom.add(0) poll
//   dsmem    st.shared::cluster into the peer CTA's shared memory, the peer polls its OWN shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_latency tools/hop_latency.cu ; run on the GPU box.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

constexpr int N = 20000;

template <int MODE>
__device__ __forceinline__ void put(unsigned int* p, unsigned int v) {
    if (MODE == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 1) *reinterpret_cast<volatile unsigned int*>(p) = v;
    if (MODE == 2) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
    if (MODE == 3) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
template <int MODE>
__device__ __forceinline__ unsigned int get(unsigned int* p) {
    unsigned int v = 0;
    if (MODE == 0) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 1) v = *reinterpret_cast<volatile unsigned int*>(p);
    if (MODE == 2) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (MODE == 3) asm volatile("atom.relaxed.gpu.global.add.u32 %0, [%1], 0;" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// CTA 0 and CTA `peer` play ping-pong; every other CTA exits (they only make sure the two land on different SMs)
template <int MODE>
__global__ void pingpong(unsigned int* a, unsigned int* b, int peer, long long* cycles) {
    if (threadIdx.x != 0) return;
    if (blockIdx.x != 0 && blockIdx.x != peer) return;
    const bool first = blockIdx.x == 0;
    unsigned int* mine = first ? a : b;
    unsigned int* theirs = first ? b : a;
    const long long t0 = clock64();
    for (unsigned int i = 1; i <= N; ++i) {
        if (first) { put<MODE>(theirs, i); while (get<MODE>(mine) < i) {} }
        else { while (get<MODE>(mine) < i) {} put<MODE>(theirs, i); }
    }
    if (first) *cycles = clock64() - t0;
}

// a chain through W warps of different CTAs: warp w waits for slot w-1 and publishes slot w (one pass = W hops)
__global__ void chain(unsigned int* slots, int rounds, long long* cycles) {
    if (threadIdx.x != 0) return;
    const int w = blockIdx.x, W = gridDim.x;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; ++r) {
        unsigned int* prev = slots + 32 * ((w + W - 1) % W);
        const unsigned int want = w == 0 ? r - 1 : r;
        while (get<0>(prev) < want) {}
        put<0>(slots + 32 * w, r);
    }
    if (w == 0) *cycles = clock64() - t0;
}

__global__ void __cluster_dims__(2, 1, 1) pingpong_dsmem(long long* cycles) {
    __shared__ unsigned int flag;
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) flag = 0;
    cl.sync();
    const unsigned int rank = cl.block_rank();
    unsigned int* remote = cl.map_shared_rank(&flag, rank ^ 1);
    if (threadIdx.x == 0) {
        volatile unsigned int* mine = &flag;
        const long long t0 = clock64();
        for (unsigned int i = 1; i <= N; ++i) {
            if (rank == 0) { *reinterpret_cast<volatile unsigned int*>(remote) = i; while (*mine < i) {} }
            else { while (*mine < i) {} *reinterpret_cast<volatile unsigned int*>(remote) = i; }
        }
        if (rank == 0) *cycles = clock64() - t0;
    }
    cl.sync();
}

int main() {
    unsigned int* flags;
    long long* cyc;
    cudaMalloc(&flags, 1 << 20);
    cudaMallocManaged(&cyc, 8);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[4] = {"gmem relaxed.gpu", "gmem volatile", "gmem release/acquire", "red / atom.add 0"};
    for (int peer : {1, 2, 75, 147}) {
        for (int mode = 0; mode < 4; ++mode) {
            cudaMemset(flags, 0, 1 << 20);
            *cyc = 0;
            unsigned int* a = flags;
            unsigned int* b = flags + 1024;
            if (mode == 0) pingpong<0><<<148, 32>>>(a, b, peer, cyc);
            if (mode == 1) pingpong<1><<<148, 32>>>(a, b, peer, cyc);
            if (mode == 2) pingpong<2><<<148, 32>>>(a, b, peer, cyc);
            if (mode == 3) pingpong<3><<<148, 32>>>(a, b, peer, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            printf("%-22s CTA0 <-> CTA%-3d  %7.1f cycles per hop  (%.0f ns)  %s\n", names[mode], peer, (double)*cyc / (2.0 * N),
                   (double)*cyc / (2.0 * N) / (clk * 1e-6), e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    for (int W : {2, 8, 64, 148}) {
        cudaMemset(flags, 0, 1 << 20);
        *cyc = 0;
        chain<<<W, 32>>>(flags, 2000, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        printf("chain over %3d CTAs                    %7.1f cycles per hop  (%.0f ns)  %s\n", W, (double)*cyc / (2000.0 * W),
               (double)*cyc / (2000.0 * W) / (clk * 1e-6), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    *cyc = 0;
    pingpong_dsmem<<<2, 32>>>(cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("dsmem (cluster of 2)                   %7.1f cycles per hop  (%.0f ns)  %s\n", (double)*cyc / (2.0 * N), (double)*cyc / (2.0 * N) / (clk * 1e-6),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    return 0;
}
