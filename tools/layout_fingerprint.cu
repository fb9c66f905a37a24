// layout_fingerprint.cu -- host-only: time and fingerprint the tile layout of csrc/sgs_tiles.cu (proposal + both sweeps) on a
// 7-point / 5-point stencil, to check that a change of the set-up code leaves every array bit for bit the same.
//   nvcc -O3 -std=c++17 -ccbin /usr/bin/g++ -Iinclude -Isparse_matrix_math_b200/csrc -gencode arch=compute_100a,code=sm_100a
//        -o tools/bin/layout_fingerprint tools/layout_fingerprint.cu && tools/bin/layout_fingerprint 128 [ny nz]
#include "../sparse_matrix_math_b200/csrc/sgs_tiles.cu"
int smm_cuda_fail(cudaError_t, const char*, const char*, int) { return 1; }
std::mutex g_smm_attr_mu;
std::atomic<long long> g_smm_launches{0};
thread_local long long t_smm_launches = 0;
thread_local bool t_smm_capturing = false;
#include <chrono>
#include <cstdio>
#include <cstdint>
static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 96, ny = argc > 2 ? atoi(argv[2]) : nx, nz = argc > 3 ? atoi(argv[3]) : nx;
    const bool wrap = argc > 4 && atoi(argv[4]) != 0;         // +-1 couplings across the grid lines: the tile graph gets cycles
    const int rows = nx * ny * nz;
    std::vector<int32_t> start(rows + 1, 0), pos, diag(rows);
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const int r = (k * ny + j) * nx + i;
        if (nz > 1 && k > 0) pos.push_back(r - nx * ny);
        if (j > 0) pos.push_back(r - nx);
        if (i > 0 || (wrap && r > 0)) pos.push_back(r - 1);
        diag[r] = (int)pos.size(); pos.push_back(r);
        if (i < nx - 1 || (wrap && r < rows - 1)) pos.push_back(r + 1);
        if (j < ny - 1) pos.push_back(r + nx);
        if (nz > 1 && k < nz - 1) pos.push_back(r + nx * ny);
        start[r + 1] = (int)pos.size();
    }
    int width = 0;
    for (int r = 0; r < rows; ++r) width = std::max(width, std::max(diag[r] - start[r], start[r + 1] - 1 - diag[r]));
    auto t0 = std::chrono::steady_clock::now();
    int ncl = 0;
    int chain_len = 1;
    std::vector<int32_t> cl = propose_grid_tiles(rows, start, pos, &ncl, &chain_len);
    if (argc > 5 && atoi(argv[5]) == 0) chain_len = 1;       // tiles in tile-level order
    auto t1 = std::chrono::steady_clock::now();
    printf("rows %d nnz %zu width %d clusters %d propose %.3f s\n", rows, pos.size(), width, ncl, std::chrono::duration<double>(t1 - t0).count());
    if (cl.empty()) return 1;
    for (int fwd = 1; fwd >= 0; --fwd) {
        SweepLayout L;
        auto a = std::chrono::steady_clock::now();
        const bool ok = layout_sweep(fwd != 0, rows, start, pos, diag, cl, ncl, width, &L, chain_len);
        auto b = std::chrono::steady_clock::now();
        uint64_t h = fnv(L.order.data(), L.order.size() * 4);
        h = fnv(L.ecol.data(), L.ecol.size() * 4, h); h = fnv(L.eidx.data(), L.eidx.size() * 4, h);
        h = fnv(L.where.data(), L.where.size() * 4, h); h = fnv(L.steps.data(), L.steps.size(), h);
        h = fnv(L.push.data(), L.push.size() * 4, h);
        // the order must be a deadlock-free schedule for warps that take whole chains in ticket order: every operand of a row
        // comes from an earlier step of its own tile, an earlier tile of its own chain, or a chain handed out before it
        bool sched = ok;
        if (ok) {
            const int clen = L.chain_len;
            for (int r = 0; r < rows && sched; ++r) {
                const int w = L.where[r], t = w >> 6;
                for (int k = fwd ? start[r] : diag[r] + 1; k < (fwd ? diag[r] : start[r + 1]) && sched; ++k) {
                    const int wo = L.where[pos[k]], to = wo >> 6;
                    if (to == t) sched = L.steps[wo] < L.steps[w];
                    else if (to / clen == t / clen) sched = to < t;
                    else sched = to / clen < t / clen;
                }
            }
        }
        printf("%s ok %d levels %d time %.3f s fingerprint %016llx chain %d schedule_ok %d\n", fwd ? "forward " : "backward", (int)ok, L.levels,
               std::chrono::duration<double>(b - a).count(), (unsigned long long)h, L.chain_len, (int)sched);
    }
    return 0;
}
