// sgs_internal.cuh -- shared between sgs.cu (row-level schedule), sgs_tiles.cu (tile-level schedule) and sgs_lines.cu (line schedule)
#pragma once
#include <stdint.h>

#include <vector>

#include "smm_internal.cuh"

struct smm_precond {
    int kind = 0;                    // 0: Symmetric Gauss-Seidel on A's values; 1: IC(0), 2: ILU(0) on their own factor values; 3: Jacobi (diagonal of A)
    float* factor = nullptr;         // IC(0): [nnz] factor in A's pattern (L below and on the diagonal, L^T above), ref H:1233-1234
                                     // ILU(0): strict L (unit diagonal implied) below, U on and above the diagonal, ref H:1203-1211
    const smm_csr* m = nullptr;
    int rows = 0;
    bool valid = true;               // structure admits the sweeps (else apply returns the reference's code 1)
    int levels_fwd = 0, levels_bwd = 0;   // row levels of the two triangles (smm_precond_levels)
    bool levels_known = false;       // computed at create time only when the row-level schedule needs them, else on demand
    long long threads_fwd = 0, threads_bwd = 0;   // padded launch sizes
    int32_t* order_fwd = nullptr;    // [threads_fwd] row index or -1 (padding)
    int32_t* order_bwd = nullptr;    // [threads_bwd]
    int32_t* diag_pos = nullptr;     // [rows] index of a_ii in positions/values
    // sliced-ELL copies of the strict lower / upper triangles in sweep order (index 0: forward, 1: backward)
    long long* slice_ptr[2] = {nullptr, nullptr};   // [threads/32 + 1]
    int32_t* ecol[2] = {nullptr, nullptr};          // column or -1 (padding)
    int32_t* eidx[2] = {nullptr, nullptr};          // index into the CSR values (to refresh eval after value updates)
    float* eval[2] = {nullptr, nullptr};
    float* dval[2] = {nullptr, nullptr};            // [threads] a_ii of the thread's row
    long long esize[2] = {0, 0};
    unsigned long long values_version = ~0ull;      // version of m->values the packed copies were gathered from
    float* yperm = nullptr;          // [threads_fwd] forward result, stored in forward sweep order
    float* xperm = nullptr;          // [threads_bwd] backward result in backward sweep order (x itself is also written in natural order)
    int32_t* ypos = nullptr;         // [threads_bwd] where the backward thread's own row sits in yperm
    unsigned int* tickets = nullptr; // [2] logical CTA counters, [2] = abort flag, [3] = error bits
    float* io[2] = {nullptr, nullptr};   // staging for the host-pointer apply
    // tile-level schedule (sgs_tiles.cu): rows grouped into tiles of up to 64, one warp per tile; when `tiled`, the
    // position-ordered arrays above are laid out tile by tile (position = 64 * tile + index inside the tile)
    bool tiled = false;
    int tile_width = 0;              // entries kept per row and sweep (<= 4), same for every tile
    uint8_t* tile_steps[2] = {nullptr, nullptr};   // [tiles * 64] step of every row, then [tiles] number of steps
    uint32_t* tile_push[2] = {nullptr, nullptr};   // [tiles * 64] operand slots inside the tile that consume the row's result
    int tile_levels[2] = {0, 0};
    std::vector<int32_t> tile_level_ptr;   // forward sweep, tiles in tile-level order: level l = tiles [ptr[l], ptr[l + 1]) (host copy; the device factorisations walk it)
    int tile_chain[2] = {1, 1};      // tiles per chain (one warp solves a chain from end to end), per sweep
    int tile_blocks = 0;             // cluster schedule: blocks of 32 chains, one thread-block cluster per block at a time (0: off)
    uint32_t* tile_push2[2] = {nullptr, nullptr};  // cluster schedule: [tiles * 64] pushes that leave the tile (next tile of the chain, other chains of the block)
    // line schedule (sgs_lines.cu): a lane per grid line, a warp per patch of 32 lines; when `lined`, yperm / xperm are ordered
    // patch by patch, step by step (position = 32 * (patch * line_steps + step) + lane), the same positions for both sweeps
    bool lined = false;
    int line_w = 0;                  // operands per row and sweep
    int line_steps = 0;              // steps per patch
    int line_patches = 0;
    int line_levels = 0;             // distinct patch offsets along j + along k - 1 (reported as the "tile levels")
    uint32_t* line_pack[2] = {nullptr, nullptr};   // [patches * steps][2 w + 2][32] per sweep
    int32_t* line_eidx[2] = {nullptr, nullptr};    // [patches * steps][w + 1][32] index of every packed coefficient / diagonal in the CSR values
};

// sgs_lines.cu
bool smm_sgs_lines_build(smm_precond* p, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos);
int smm_sgs_lines_gather(const smm_precond* p, cudaStream_t s);
int smm_sgs_lines_launch(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, unsigned int sleep_first, unsigned int sleep_later, cudaStream_t s);

// sgs_tiles.cu
bool smm_sgs_detect_grid(int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, long long* nx, long long* ny, long long* nz);
bool smm_sgs_grid_from_offsets(int rows, std::vector<long long> offs, long long* nx, long long* ny, long long* nz);
bool smm_sgs_tiles_build(smm_precond* p, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, const std::vector<int32_t>& diag);
int smm_sgs_tiles_launch(const smm_precond* p, const float* rhs_dev, float* x_dev, SolveState* state, int ctas_per_sm, unsigned int sleep_first,
                         unsigned int sleep_later, cudaStream_t s);

// sgs_tiles_setup.cu: the same tile layout built by kernels from the CSR arrays in HBM (no download, no upload).
// diag_dev: [rows] index of a_ii (device).  Returns false when the device path does not apply or does not verify -- the
// caller then runs smm_sgs_tiles_build on host copies, which decides for good.
bool smm_sgs_tiles_build_dev(smm_precond* p, const smm_csr* m, const int32_t* diag_dev, int width);
// diagonal positions and structural validity (find_diagonals of sgs.cu) on the device: *valid, *width = most entries a row
// keeps on one side of its diagonal
int smm_sgs_diagonals_dev(const smm_csr* m, int32_t* diag_dev, bool* valid, int* width);

// IC(0) (kind 1) / ILU(0) (kind 2) factorisation on the device, in the order of the forward tile schedule: needs p->tiled with
// single-tile chains and p->tile_level_ptr.  *code: 0 ok, 2 = ILU(0) pivot not > 1e-6 in magnitude; returns SMM_E_STATE when
// the result must come from the host code instead (no tile schedule, a non-finite IC(0) pivot).
int smm_sgs_factorize_dev(smm_precond* p, int* code);
void smm_sgs_tiles_release(smm_precond* p);   // drops a tile schedule built by either path

namespace {

constexpr int TILE = 64;         // rows per tile (sgs_tiles.cu)
constexpr int TILE_MAX_W = 4;    // stored operands per row and sweep
constexpr int MAX_STEPS = 64;    // internal levels of a tile
constexpr int MAX_PREDS = 8;     // distinct predecessor tiles of a tile

// proposed tile of a natural-order grid: 4 x 4 x 4 grid points in 3D, 8 x 8 in 2D
inline void smm_sgs_tile_shape(long long nz, int* ti, int* tj, int* tk) {
    *ti = nz > 1 ? 4 : 8; *tj = nz > 1 ? 4 : 8; *tk = nz > 1 ? 4 : 1;
}

constexpr unsigned int SENTINEL = 0x7FC0DEADu;   // quiet NaN with a payload; GPU arithmetic only produces 0x7FFFFFFF
constexpr int SGS_THREADS = 128;
constexpr unsigned int POLL_LIMIT = 1u << 22;

__device__ __forceinline__ unsigned int peek(const float* p) {
    unsigned int bits;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(bits) : "l"(p));
    return bits;
}

// back off between polls: thousands of lanes may be waiting on L2.  Returns false when the wait must be abandoned.
__device__ __forceinline__ bool poll_pause(unsigned int* polls, unsigned int* abort_flag, unsigned int sleep_first, unsigned int sleep_later) {
    const unsigned int ns = *polls < 16u ? sleep_first : sleep_later;
    if (ns) __nanosleep(ns);
    if ((++*polls & 255u) == 0u) {
        if (peek(reinterpret_cast<const float*>(abort_flag)) != 0u || *polls >= POLL_LIMIT) { atomicExch(abort_flag, 1u); return false; }
    }
    return true;
}

__device__ __forceinline__ void publish(float* p, float v) {
    // a computed value can never equal the sentinel payload, so the store itself is the ready flag
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace
