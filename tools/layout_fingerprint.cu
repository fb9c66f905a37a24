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
#include <cmath>
#include <cstring>
#include <limits>
#include <cstdio>
#include <cstdint>
static uint64_t fnv(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Host emulation of sgs_cluster_kernel's data flow, driven by nothing but the layout arrays: every operand is taken from
// where the kernel would take it (the warp's double-buffered staging, its inbox, or the published vector), every result goes
// where the kernel would push it; a row whose operand has not arrived blocks its tile, and chains are advanced round-robin
// until nothing moves.  The result must equal a plain triangular solve bit for bit, with every row solved (no deadlock).
static bool emulate_cluster_sweep(bool fwd, int rows, const std::vector<int32_t>& start, const std::vector<int32_t>& pos, const std::vector<int32_t>& diag,
                                  int width, const SweepLayout& L) {
    if (L.nblocks <= 0) return false;
    const int clen = L.chain_len;
    const long long ntiles = L.ntiles, nchains = ntiles / clen;
    auto val = [&](int k) { return -1.0f - 0.03125f * (float)(k % 7); };          // off-diagonal a_k; diagonal 8
    std::vector<float> rhs(rows), ref(rows);
    for (int r = 0; r < rows; ++r) rhs[r] = 1.0f + (float)((r * 2654435761u) >> 20) / 4096.0f;
    for (int q = 0; q < rows; ++q) {                                             // the reference order of the sweep
        const int r = fwd ? q : rows - 1 - q;
        float acc = rhs[r];
        if (fwd) for (int k = start[r]; k < diag[r]; ++k) acc = acc - val(k) * ref[pos[k]];
        else for (int k = start[r + 1] - 1; k > diag[r]; --k) acc = acc - val(k) * ref[pos[k]];
        ref[r] = acc / 8.0f;
    }
    const float NANV = std::numeric_limits<float>::quiet_NaN();
    std::vector<float> glob((size_t)ntiles * TILE, NANV), stage((size_t)nchains * 2 * 256, NANV), inbox((size_t)nchains * clen * INBOX_SLOTS, NANV);
    std::vector<int> at(nchains, 0), row_at(nchains, 0);                        // tile of the chain, next row (in step order) inside it
    std::vector<std::vector<int>> by_step(ntiles);
    for (long long t = 0; t < ntiles; ++t) {
        for (int i = 0; i < TILE; ++i) if (L.order[t * TILE + i] >= 0) by_step[t].push_back(i);
        std::stable_sort(by_step[t].begin(), by_step[t].end(), [&](int x, int y) { return L.steps[t * TILE + x] < L.steps[t * TILE + y]; });
    }
    std::vector<float> got(rows, NANV);
    long long solved = 0;
    for (bool moved = true; moved;) {
        moved = false;
        for (long long c = 0; c < nchains; ++c) {
            while (at[c] < clen) {
                const long long t = c * clen + at[c];
                float* mine = &stage[(size_t)c * 512 + (size_t)(at[c] & 1) * 256];
                float* next = &stage[(size_t)c * 512 + (size_t)((at[c] & 1) ^ 1) * 256];
                if (row_at[c] == 0) {                                             // tile start: the slots nobody pushes into
                    for (int i = 0; i < TILE; ++i) for (int e = 0; e < TILE_MAX_W; ++e) {
                        const int code = e < width ? L.ecol[((size_t)t * width + e) * TILE + i] : E_NONE;
                        if (code == E_NONE) mine[4 * i + e] = 0.0f;
                    }
                }
                bool blocked = false;
                while (row_at[c] < (int)by_step[t].size()) {
                    const int i = by_step[t][row_at[c]], r = L.order[t * TILE + i];
                    float o[TILE_MAX_W], a[TILE_MAX_W];
                    for (int e = 0; e < TILE_MAX_W && !blocked; ++e) {
                        const int code = e < width ? L.ecol[((size_t)t * width + e) * TILE + i] : E_NONE;
                        const int k = e < width ? L.eidx[((size_t)t * width + e) * TILE + i] : -1;
                        a[e] = k >= 0 ? val(k) : 0.0f;
                        if (code >= 0) o[e] = glob[code];
                        else if (code <= E_INBOX) o[e] = inbox[((size_t)c * clen + at[c]) * INBOX_SLOTS + (E_INBOX - code)];
                        else o[e] = mine[4 * i + e];                              // E_LOCAL (pushed) or E_NONE (zeroed above)
                        if (std::isnan(o[e])) blocked = true;
                    }
                    if (blocked) break;
                    float acc = rhs[r];
                    for (int e = 0; e < TILE_MAX_W; ++e) acc = acc - a[e] * o[e];
                    const float res = acc / 8.0f;
                    const uint32_t pu = L.push[t * TILE + i], p2 = L.push2[t * TILE + i];
                    mine[pu & 255u] = res; mine[(pu >> 8) & 255u] = res; mine[(pu >> 16) & 255u] = res;
                    if ((p2 & 0xFFu) != 0xFFu) next[p2 & 0xFFu] = res;
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t rr = (p2 >> (8 + 11 * j)) & 0x7FFu;
                        if (rr != 0x7FFu) inbox[(((size_t)(c / CLUSTER_CHAINS) * CLUSTER_CHAINS + (rr >> 5)) * clen + at[c]) * INBOX_SLOTS + (rr & 31u)] = res;
                    }
                    glob[t * TILE + i] = res;
                    got[r] = res;
                    ++solved; ++row_at[c];
                    moved = true;
                }
                if (blocked) break;
                // tile finished: the staging buffer it used is recycled by tile at + 2 (its pushed slots are rewritten by then)
                for (int q = 0; q < 256; ++q) mine[q] = NANV;
                ++at[c]; row_at[c] = 0;
                moved = true;
            }
        }
    }
    if (solved != rows) {
        printf("emulation: %lld of %d rows solved (deadlock)\n", solved, rows);
        int shown = 0;
        for (long long c = 0; c < nchains && shown < 6; ++c) {
            if (at[c] >= clen) continue;
            const long long t = c * clen + at[c];
            if (row_at[c] >= (int)by_step[t].size()) continue;
            const int i = by_step[t][row_at[c]], r = L.order[t * TILE + i];
            printf("  chain %lld (block %lld warp %lld) tile at %d row slot %d (row %d) waits:", c, c / CLUSTER_CHAINS, c % CLUSTER_CHAINS, at[c], i, r);
            for (int e = 0; e < width; ++e) {
                const int code = L.ecol[((size_t)t * width + e) * TILE + i], k = L.eidx[((size_t)t * width + e) * TILE + i];
                if (k < 0) continue;
                const int w = L.where[pos[k]];
                printf(" [e%d code %d from row %d at tile %d (chain %d at %d) slot %d push2 %08x]", e, code, pos[k], w >> 6, (w >> 6) / clen, (w >> 6) % clen, w & 63, L.push2[w]);
            }
            printf("\n");
            ++shown;
        }
        return false;
    }
    for (int r = 0; r < rows; ++r) if (memcmp(&got[r], &ref[r], 4) != 0) { printf("emulation: row %d differs (%g vs %g)\n", r, got[r], ref[r]); return false; }
    return true;
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
    ClusterPlan plan;
    std::vector<int32_t> cl = propose_grid_tiles(rows, start, pos, &ncl, &chain_len, &plan);
    if (argc > 5 && atoi(argv[5]) == 0) chain_len = 1;       // tiles in tile-level order
    const bool clusters = argc > 6 && atoi(argv[6]) != 0;    // cluster schedule: laid out, then EMULATED on the host (see below)
    auto t1 = std::chrono::steady_clock::now();
    printf("rows %d nnz %zu width %d clusters %d propose %.3f s\n", rows, pos.size(), width, ncl, std::chrono::duration<double>(t1 - t0).count());
    if (cl.empty()) return 1;
    for (int fwd = 1; fwd >= 0; --fwd) {
        SweepLayout L;
        auto a = std::chrono::steady_clock::now();
        const bool ok = layout_sweep(fwd != 0, rows, start, pos, diag, cl, ncl, width, &L, chain_len, clusters ? &plan : nullptr);
        auto b = std::chrono::steady_clock::now();
        uint64_t h = fnv(L.order.data(), L.order.size() * 4);
        h = fnv(L.ecol.data(), L.ecol.size() * 4, h); h = fnv(L.eidx.data(), L.eidx.size() * 4, h);
        h = fnv(L.where.data(), L.where.size() * 4, h); h = fnv(L.steps.data(), L.steps.size(), h);
        h = fnv(L.push.data(), L.push.size() * 4, h);
        // the order must be a deadlock-free schedule for warps that take whole chains in ticket order: every operand of a row
        // comes from an earlier step of its own tile, an earlier tile of its own chain, or a chain handed out before it
        if (ok && clusters) {
            h = fnv(L.push2.data(), L.push2.size() * 4, h);
            printf("%s cluster blocks %d tiles %lld emulation_ok %d\n", fwd ? "forward " : "backward", L.nblocks, L.ntiles,
                   (int)emulate_cluster_sweep(fwd != 0, rows, start, pos, diag, width, L));
        }
        bool sched = ok;
        if (ok) {
            const int clen = L.chain_len;
            for (int r = 0; r < rows && sched; ++r) {
                const int w = L.where[r], t = w >> 6;
                for (int k = fwd ? start[r] : diag[r] + 1; k < (fwd ? diag[r] : start[r + 1]) && sched; ++k) {
                    const int wo = L.where[pos[k]], to = wo >> 6;
                    if (to == t) sched = L.steps[wo] < L.steps[w];
                    else if (to / clen == t / clen) sched = to < t;
                    else if (L.nblocks > 0) sched = to / clen / CLUSTER_CHAINS <= t / clen / CLUSTER_CHAINS;   // same block (concurrent warps) or a block handed out earlier
                    else sched = to / clen < t / clen;
                }
            }
        }
        printf("%s ok %d levels %d time %.3f s fingerprint %016llx chain %d schedule_ok %d\n", fwd ? "forward " : "backward", (int)ok, L.levels,
               std::chrono::duration<double>(b - a).count(), (unsigned long long)h, L.chain_len, (int)sched);
    }
    return 0;
}
