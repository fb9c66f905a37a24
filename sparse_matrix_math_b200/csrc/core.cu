// core.cu -- library state, error reporting, CSR handles, workspaces, host-pointer wrappers of SpMV and dot.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "epilogue.cuh"
#include "smm_internal.cuh"

std::atomic<long long> g_smm_launches{0};
thread_local long long t_smm_launches = 0;
thread_local bool t_smm_capturing = false;
std::mutex g_smm_attr_mu;

// Measured (profiles/r02_pdl.txt): with plain stream launches the programmatic dependency takes 4-7 us off an L2-resident CG
// iteration (512^2: 22.6 -> 15.9 us, 1024^2: 28.8 -> 24.4 us); inside the captured iteration graphs it costs 1-4 % at every
// size (512^3: 464 -> 444 it/s).  Default: on for launches outside a stream capture, off inside one.  SMM_B200_PDL=1 / 0
// forces it on / off everywhere.
bool smm_pdl_enabled() {
    static const int mode = [] { const char* e = getenv("SMM_B200_PDL"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    return mode >= 0 ? mode == 1 : !t_smm_capturing;
}

namespace {
thread_local char g_err[512] = "";
cudaStream_t g_streams[64] = {nullptr};
std::mutex g_mu;
}  // namespace

void smm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int smm_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    smm_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return SMM_E_CUDA;
}

cudaStream_t smm_default_stream() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_streams[dev]) {
        if (cudaStreamCreateWithFlags(&g_streams[dev], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    }
    return g_streams[dev];
}

static inline cudaStream_t pick(void* stream) { return stream ? (cudaStream_t)stream : smm_default_stream(); }

// ---------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------
int smm_workspace_get(const smm_csr* mc, smm_workspace** out) {
    smm_csr* m = const_cast<smm_csr*>(mc);
    if (!m->ws) {
        smm_workspace* ws = new smm_workspace();
        ws->device = m->device;
        const int rc = [&]() -> int {                          // a workspace is attached whole or not at all
            cudaDeviceProp prop;
            SMM_CUDA(cudaGetDeviceProperties(&prop, m->device));
            ws->sm_count = prop.multiProcessorCount;
            size_t cap = (size_t)m->num_blocks;
            const size_t vg = (size_t)smm_vec_max_grid(ws);
            if (cap < vg) cap = vg;
            cap = (cap + 63) & ~(size_t)63;
            ws->partials_cap = cap;
            SMM_CUDA(cudaMalloc(&ws->partials, sizeof(float) * cap * 2 * RED_SLOTS));
            SMM_CUDA(cudaMalloc(&ws->tickets, sizeof(unsigned int) * RED_SLOTS));
            SMM_CUDA(cudaMemset(ws->tickets, 0, sizeof(unsigned int) * RED_SLOTS));
            SMM_CUDA(cudaMalloc(&ws->grid_barrier, sizeof(unsigned int) * 2));
            SMM_CUDA(cudaMemset(ws->grid_barrier, 0, sizeof(unsigned int) * 2));
            SMM_CUDA(cudaMalloc(&ws->state, sizeof(SolveState)));
            SMM_CUDA(cudaMemset(ws->state, 0, sizeof(SolveState)));
            SMM_CUDA(cudaMallocHost(&ws->state_host, sizeof(SolveState)));
            SMM_CUDA(cudaEventCreate(&ws->ev0));
            SMM_CUDA(cudaEventCreate(&ws->ev1));
            return SMM_OK;
        }();
        if (rc != SMM_OK) { smm_workspace_free(ws); return rc; }
        m->ws = ws;
    }
    *out = m->ws;
    return SMM_OK;
}

int smm_workspace_vectors(smm_workspace* ws, int count, size_t len) {
    if (count > 10) return SMM_E_INVALID;
    if (len == 0) len = 1;
    if (ws->vec_len < len) {
        for (int i = 0; i < 10; ++i) if (ws->vec[i]) { cudaFree(ws->vec[i]); ws->vec[i] = nullptr; }
        ws->vec_len = len;
    }
    for (int i = 0; i < count; ++i) {
        if (!ws->vec[i]) SMM_CUDA(cudaMalloc(&ws->vec[i], sizeof(float) * ((ws->vec_len + 3) & ~(size_t)3)));
    }
    return SMM_OK;
}

void smm_workspace_free(smm_workspace* ws) {
    if (!ws) return;
    cudaFree(ws->partials);
    cudaFree(ws->tickets);
    cudaFree(ws->grid_barrier);
    cudaFree(ws->state);
    cudaFreeHost(ws->state_host);
    cudaFree(ws->history);
    smm_dot_scratch_free(&ws->dot);
    for (int i = 0; i < 10; ++i) cudaFree(ws->vec[i]);
    if (ws->ev0) cudaEventDestroy(ws->ev0);
    if (ws->ev1) cudaEventDestroy(ws->ev1);
    delete ws;
}

// ---------------------------------------------------------------------------------------------------
// C ABI: library
// ---------------------------------------------------------------------------------------------------
extern "C" {

int smm_abi_version(void) { return SMM_B200_ABI_VERSION; }
const char* smm_last_error(void) { return g_err; }
long long smm_kernel_launch_count(void) { return g_smm_launches.load(); }

int smm_device_count(int* count) {
    SMM_CUDA(cudaGetDeviceCount(count));
    return SMM_OK;
}

int smm_set_device(int device) {
    SMM_CUDA(cudaSetDevice(device));
    return SMM_OK;
}

int smm_device_info(int* sm_count, size_t* l2_bytes, size_t* total_mem, size_t* free_mem) {
    int dev = 0;
    SMM_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SMM_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    size_t f = 0, t = 0;
    SMM_CUDA(cudaMemGetInfo(&f, &t));
    if (total_mem) *total_mem = t;
    if (free_mem) *free_mem = f;
    return SMM_OK;
}

int smm_sync(void) {
    SMM_CUDA(cudaDeviceSynchronize());
    return SMM_OK;
}

int smm_malloc_dev(size_t bytes, void** p) {
    SMM_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return SMM_OK;
}
int smm_free_dev(void* p) {
    SMM_CUDA(cudaFree(p));
    return SMM_OK;
}
int smm_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    SMM_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return SMM_OK;
}
int smm_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    // the library's streams are non-blocking: wait for everything in flight before the (legacy-stream) copy
    SMM_CUDA(cudaDeviceSynchronize());
    SMM_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return SMM_OK;
}
int smm_memset_dev(void* dst, int byte, size_t bytes) {
    SMM_CUDA(cudaMemset(dst, byte, bytes));
    return SMM_OK;
}

// ---------------------------------------------------------------------------------------------------
// CSR handles
// ---------------------------------------------------------------------------------------------------
static int csr_finish_create(smm_csr* m) {
    cudaStream_t s = smm_default_stream();
    SMM_TRY(smm_csr_analyse(m, s));
    SMM_TRY(smm_first_active_start(m, &m->first_active_start, s));
    return SMM_OK;
}

int smm_csr_create(int rows, int cols, const int32_t* start, const int32_t* positions, const float* values, smm_csr_t** out) {
    if (!out || rows < 0 || cols < 0 || (rows > 0 && !start)) { smm_set_error("smm_csr_create: bad arguments"); return SMM_E_INVALID; }
    smm_csr* m = new smm_csr();
    if (cudaGetDevice(&m->device) != cudaSuccess) { delete m; return smm_cuda_fail(cudaGetLastError(), "cudaGetDevice", __FILE__, __LINE__); }
    m->rows = rows; m->cols = cols;
    m->nnz = start ? start[rows] : 0;
    if (m->nnz < 0 || (m->nnz > 0 && (!positions || !values))) { delete m; smm_set_error("smm_csr_create: bad arrays"); return SMM_E_INVALID; }
    const size_t npad = ((size_t)m->nnz + 3) & ~(size_t)3;
    m->nnz_alloc = (int64_t)(npad ? npad : 4);
    int rc = SMM_OK;
    do {
        if (cudaMalloc(&m->start, sizeof(int32_t) * ((size_t)rows + 1)) != cudaSuccess ||
            cudaMalloc(&m->positions, sizeof(int32_t) * (npad ? npad : 4)) != cudaSuccess ||
            cudaMalloc(&m->values, sizeof(float) * (npad ? npad : 4)) != cudaSuccess) { rc = smm_cuda_fail(cudaGetLastError(), "cudaMalloc(csr)", __FILE__, __LINE__); break; }
        if (start) {
            if (cudaMemcpy(m->start, start, sizeof(int32_t) * ((size_t)rows + 1), cudaMemcpyHostToDevice) != cudaSuccess) { rc = smm_cuda_fail(cudaGetLastError(), "memcpy(start)", __FILE__, __LINE__); break; }
        } else {
            cudaMemset(m->start, 0, sizeof(int32_t));
        }
        if (m->nnz) {
            if (cudaMemcpy(m->positions, positions, sizeof(int32_t) * (size_t)m->nnz, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaMemcpy(m->values, values, sizeof(float) * (size_t)m->nnz, cudaMemcpyHostToDevice) != cudaSuccess) { rc = smm_cuda_fail(cudaGetLastError(), "memcpy(csr)", __FILE__, __LINE__); break; }
        }
        rc = csr_finish_create(m);
    } while (0);
    if (rc != SMM_OK) { smm_csr_destroy(m); return rc; }
    *out = m;
    return SMM_OK;
}

int smm_csr_create_dev(int rows, int cols, int32_t* start_dev, int32_t* positions_dev, float* values_dev, int copy, smm_csr_t** out) {
    if (!out || rows < 0 || cols < 0 || !start_dev) { smm_set_error("smm_csr_create_dev: bad arguments"); return SMM_E_INVALID; }
    // Ownership: with copy == 0 the caller's arrays become the handle's ONLY when the call succeeds; on any failure they are
    // left untouched and still belong to the caller.  Everything this function allocated itself is released on failure.
    smm_csr* m = new smm_csr();
    m->rows = rows; m->cols = cols;
    int rc = SMM_OK;
    int32_t nnz32 = 0;
    do {
        if (cudaGetDevice(&m->device) != cudaSuccess ||
            cudaMemcpy(&nnz32, start_dev + rows, sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = smm_cuda_fail(cudaGetLastError(), "smm_csr_create_dev: read nnz", __FILE__, __LINE__); break; }
        m->nnz = nnz32;
        m->nnz_alloc = m->nnz;                                 // adopted arrays: nothing is known beyond nnz entries
        if (m->nnz < 0 || (m->nnz > 0 && (!positions_dev || !values_dev))) { smm_set_error("smm_csr_create_dev: bad arrays"); rc = SMM_E_INVALID; break; }
        if (copy) {
            const size_t npad = ((size_t)m->nnz + 3) & ~(size_t)3;
            m->nnz_alloc = (int64_t)(npad ? npad : 4);
            if (cudaMalloc(&m->start, sizeof(int32_t) * ((size_t)rows + 1)) != cudaSuccess ||
                cudaMalloc(&m->positions, sizeof(int32_t) * (npad ? npad : 4)) != cudaSuccess ||
                cudaMalloc(&m->values, sizeof(float) * (npad ? npad : 4)) != cudaSuccess ||
                cudaMemcpy(m->start, start_dev, sizeof(int32_t) * ((size_t)rows + 1), cudaMemcpyDeviceToDevice) != cudaSuccess ||
                (m->nnz && cudaMemcpy(m->positions, positions_dev, sizeof(int32_t) * (size_t)m->nnz, cudaMemcpyDeviceToDevice) != cudaSuccess) ||
                (m->nnz && cudaMemcpy(m->values, values_dev, sizeof(float) * (size_t)m->nnz, cudaMemcpyDeviceToDevice) != cudaSuccess)) {
                rc = smm_cuda_fail(cudaGetLastError(), "smm_csr_create_dev: copy", __FILE__, __LINE__);
                break;
            }
        } else {
            if (((uintptr_t)positions_dev & 15) || ((uintptr_t)values_dev & 15)) { smm_set_error("smm_csr_create_dev: adopted arrays must be 16-byte aligned"); rc = SMM_E_INVALID; break; }
            m->start = start_dev; m->positions = positions_dev; m->values = values_dev;
        }
        m->owns_arrays = true;
        rc = csr_finish_create(m);
    } while (0);
    if (rc != SMM_OK) {
        if (!copy) { m->start = nullptr; m->positions = nullptr; m->values = nullptr; }   // still the caller's
        smm_csr_destroy(m);
        return rc;
    }
    *out = m;
    return SMM_OK;
}

int smm_csr_update_values(smm_csr_t* m, const float* values) {
    if (!m || (m->nnz && !values)) return SMM_E_INVALID;
    SMM_CUDA(cudaSetDevice(m->device));
    if (m->nnz) SMM_CUDA(cudaMemcpy(m->values, values, sizeof(float) * (size_t)m->nnz, cudaMemcpyHostToDevice));
    m->values_version++;
    return SMM_OK;
}

int smm_csr_destroy(smm_csr_t* m) {
    if (!m) return SMM_OK;
    cudaSetDevice(m->device);
    if (m->owns_arrays) { cudaFree(m->start); cudaFree(m->positions); cudaFree(m->values); }
    cudaFree(m->block_row);
    cudaFree(m->row_perm);
    cudaFree(m->start16);
    smm_workspace_free(m->ws);
    delete m;
    return SMM_OK;
}

int smm_csr_shape(const smm_csr_t* m, int* rows, int* cols, int64_t* nnz, int* first_active_start) {
    if (!m) return SMM_E_INVALID;
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (nnz) *nnz = m->nnz;
    if (first_active_start) *first_active_start = m->first_active_start;
    return SMM_OK;
}

int smm_csr_download(const smm_csr_t* m, int32_t* start, int32_t* positions, float* values) {
    if (!m) return SMM_E_INVALID;
    SMM_CUDA(cudaSetDevice(m->device));
    if (start) SMM_CUDA(cudaMemcpy(start, m->start, sizeof(int32_t) * ((size_t)m->rows + 1), cudaMemcpyDeviceToHost));
    if (positions && m->nnz) SMM_CUDA(cudaMemcpy(positions, m->positions, sizeof(int32_t) * (size_t)m->nnz, cudaMemcpyDeviceToHost));
    if (values && m->nnz) SMM_CUDA(cudaMemcpy(values, m->values, sizeof(float) * (size_t)m->nnz, cudaMemcpyDeviceToHost));
    return SMM_OK;
}

int smm_csr_device_arrays(const smm_csr_t* m, const int32_t** start_dev, const int32_t** positions_dev, const float** values_dev) {
    if (!m) return SMM_E_INVALID;
    if (start_dev) *start_dev = m->start;
    if (positions_dev) *positions_dev = m->positions;
    if (values_dev) *values_dev = m->values;
    return SMM_OK;
}

// ---------------------------------------------------------------------------------------------------
// SpMV
// ---------------------------------------------------------------------------------------------------
int smm_spmv_dev(const smm_csr_t* m, int op, const float* lhs_dev, const float* mult_dev, float* out_dev, int exact, void* stream) {
    if (!m || op < 0 || op > 2 || (m->rows && !out_dev) || (m->nnz && !mult_dev) || (op != SMM_OP_ASSIGN && m->rows && !lhs_dev)) {
        smm_set_error("smm_spmv: bad arguments");
        return SMM_E_INVALID;
    }
    if (mult_dev == out_dev && m->rows) { smm_set_error("smm_spmv: mult must not alias out (H:1503)"); return SMM_E_ALIAS; }
    SpmvArgs a;
    a.m = m; a.op = op; a.lhs = lhs_dev; a.mult = mult_dev; a.out = out_dev; a.exact = exact;
    return smm_launch_spmv(a, pick(stream));
}

int smm_spmv(const smm_csr_t* m, int op, const float* lhs, const float* mult, float* out) {
    if (!m || op < 0 || op > 2) return SMM_E_INVALID;
    if (m->rows == 0) return SMM_OK;
    if (!out || (m->cols && !mult) || (op != SMM_OP_ASSIGN && !lhs)) { smm_set_error("smm_spmv: null vector"); return SMM_E_INVALID; }
    if (mult == out) { smm_set_error("smm_spmv: mult must not alias out (H:1503)"); return SMM_E_ALIAS; }
    SMM_CUDA(cudaSetDevice(m->device));
    smm_workspace* ws = nullptr;
    SMM_TRY(smm_workspace_get(m, &ws));
    const size_t len = (size_t)(m->rows > m->cols ? m->rows : m->cols);
    SMM_TRY(smm_workspace_vectors(ws, 3, len));
    cudaStream_t s = smm_default_stream();
    float *d_lhs = ws->vec[0], *d_mult = ws->vec[1], *d_out = ws->vec[2];
    if (op != SMM_OP_ASSIGN) SMM_CUDA(cudaMemcpyAsync(d_lhs, lhs, sizeof(float) * (size_t)m->rows, cudaMemcpyHostToDevice, s));
    if (m->cols) SMM_CUDA(cudaMemcpyAsync(d_mult, mult, sizeof(float) * (size_t)m->cols, cudaMemcpyHostToDevice, s));
    SMM_TRY(smm_spmv_dev(m, op, d_lhs, d_mult, d_out, 0, s));
    SMM_CUDA(cudaMemcpyAsync(out, d_out, sizeof(float) * (size_t)m->rows, cudaMemcpyDeviceToHost, s));
    SMM_CUDA(cudaStreamSynchronize(s));
    return SMM_OK;
}

}  // extern "C"
