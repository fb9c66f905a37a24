// factor_fingerprint.cu -- host-only: the IC(0) and ILU(0) factorisations of csrc/sgs.cu (set-up code that runs on the
// host) on generated stencils, fingerprinted so that tests/test_tile_layout_cpu.py can compare them with the oracle's
// factors bit for bit without a GPU.
//   nvcc -O2 -std=c++17 -ccbin /usr/bin/g++ -Iinclude -Isparse_matrix_math_b200/csrc -gencode arch=compute_100a,code=sm_100a
//        -o tools/bin/factor_fingerprint tools/factor_fingerprint.cu && tools/bin/factor_fingerprint nx ny nz c
#include "../sparse_matrix_math_b200/csrc/sgs.cu"
int smm_cuda_fail(cudaError_t, const char*, const char*, int) { return 1; }
void smm_set_error(const char*, ...) {}
cudaStream_t smm_default_stream() { return nullptr; }
bool smm_sgs_tiles_build(smm_precond*, int, const std::vector<int32_t>&, const std::vector<int32_t>&, const std::vector<int32_t>&) { return false; }
int smm_sgs_tiles_launch(const smm_precond*, const float*, float*, SolveState*, int, unsigned int, unsigned int, cudaStream_t) { return 0; }
bool smm_sgs_tiles_build_dev(smm_precond*, const smm_csr*, const int32_t*, int) { return false; }
int smm_sgs_diagonals_dev(const smm_csr*, int32_t*, bool*, int*) { return 1; }
int smm_sgs_factorize_dev(smm_precond*, int*) { return 1; }
void smm_sgs_tiles_release(smm_precond*) {}
bool smm_sgs_lines_build(smm_precond*, int, const std::vector<int32_t>&, const std::vector<int32_t>&) { return false; }
int smm_sgs_lines_gather(const smm_precond*, cudaStream_t) { return 0; }
int smm_sgs_lines_launch(const smm_precond*, const float*, float*, SolveState*, unsigned int, unsigned int, cudaStream_t) { return 0; }
std::atomic<long long> g_smm_launches{0};
thread_local long long t_smm_launches = 0;
thread_local bool t_smm_capturing = false;
std::mutex g_smm_attr_mu;
#include <cstdint>
#include <cstdio>
static uint64_t fnv(const void* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
int main(int argc, char** argv) {
    const int nx = argc > 1 ? atoi(argv[1]) : 12, ny = argc > 2 ? atoi(argv[2]) : nx, nz = argc > 3 ? atoi(argv[3]) : nx;
    const float c = argc > 4 ? (float)atof(argv[4]) : 0.5f;
    const float lo = -1.0f - c, hi = -1.0f + c, dg = nz > 1 ? 6.0f : 4.0f;      // tests/matgen.py: convdiff3d / poisson2d
    const int rows = nx * ny * nz;
    std::vector<int32_t> start(rows + 1, 0), pos, diag(rows);
    std::vector<float> val;
    for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i) {
        const int r = (k * ny + j) * nx + i;
        if (nz > 1 && k > 0) { pos.push_back(r - nx * ny); val.push_back(lo); }
        if (j > 0) { pos.push_back(r - nx); val.push_back(lo); }
        if (i > 0) { pos.push_back(r - 1); val.push_back(lo); }
        diag[r] = (int)pos.size(); pos.push_back(r); val.push_back(dg);
        if (i < nx - 1) { pos.push_back(r + 1); val.push_back(hi); }
        if (j < ny - 1) { pos.push_back(r + nx); val.push_back(hi); }
        if (nz > 1 && k < nz - 1) { pos.push_back(r + nx * ny); val.push_back(hi); }
        start[r + 1] = (int)pos.size();
    }
    std::vector<int32_t> d2;
    const bool valid = find_diagonals(rows, start, pos, 0, &d2);
    std::vector<float> ic0, ilu0;
    const int rc_ic0 = ic0_factorize_host(rows, start, pos, d2, val, &ic0);
    const int rc_ilu0 = ilu0_factorize_host(rows, start, pos, d2, val, &ilu0);
    std::vector<int32_t> of, ob;
    int lf = 0, lb = 0;
    std::vector<int32_t> lev_f, lev_b;
    row_levels(rows, start, pos, d2, &lev_f, &lev_b, &lf, &lb);    // the row-level schedule's analysis
    level_orders(rows, lev_f, lev_b, lf, lb, &of, &ob);
    bool perm = of.size() % 32 == 0 && ob.size() % 32 == 0;
    {
        std::vector<char> seen(rows, 0);
        for (int32_t r : of) if (r >= 0) { if (seen[r]) perm = false; seen[r] = 1; }
        for (char c2 : seen) if (!c2) perm = false;
    }
    printf("levels %d %d order_is_permutation %d\n", lf, lb, (int)perm);
    printf("rows %d nnz %zu valid %d diag %d ic0 rc %d %016llx ilu0 rc %d %016llx\n", rows, pos.size(), (int)valid, (int)(d2 == diag), rc_ic0,
           (unsigned long long)fnv(ic0.data(), ic0.size() * 4), rc_ilu0, (unsigned long long)fnv(ilu0.data(), ilu0.size() * 4));
    return 0;
}
