/*
 * oracle/smm_oracle.c -- TEST INFRASTRUCTURE ONLY (see smm_oracle.h).
 *
 * Plain-C restatement of the float Krylov hot path of vasil-pashov/sparse_matrix_math.
 * "H:n" = line n of the reference's include/sparse_matrix_math.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off keeps `a*x+b` at two roundings, the reference's default `_smm_fma` (H:27-37).
 *
 * OpenMP only parallelises loops whose result is independent of the schedule (row loops, element-wise
 * loops, the leaves of the deterministic reduce tree), exactly the loops the reference hands to TBB.
 */
#include "smm_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * _smm_fma, H:27-37 (default branch: a * x + b, two roundings)
 * ---------------------------------------------------------------------------------------------- */
static inline float smm_fma(float a, float x, float b) { return a * x + b; }

int smm_oracle_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void smm_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void smm_oracle_free(void *p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * TripletMatrix::addEntry (H:606-618) + CSRMatrix::fillArrays (H:1606-1641)
 *
 * std::map<uint64_t,T> keyed (row<<32)|col: the first addEntry stores the value, later ones `+=` it in
 * call order; iteration is in ascending key order.  Restated as: stable sort of the call sequence by
 * key, then a left-to-right float sum inside each run of equal keys.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { uint64_t key; int64_t seq; } trip_key;

static void trip_merge_sort(trip_key *a, trip_key *tmp, int64_t n) {
    /* bottom-up stable merge sort on key (seq breaks no ties: stability keeps call order) */
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t i = lo, j = mid, k = lo;
            while (i < mid && j < hi) tmp[k++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
            while (i < mid) tmp[k++] = a[i++];
            while (j < hi) tmp[k++] = a[j++];
        }
        memcpy(a, tmp, (size_t)n * sizeof(trip_key));
    }
}

int smm_oracle_triplets_to_csr(int rows, int cols, int64_t n_triplets,
                               const int *trow, const int *tcol, const float *tval,
                               int *start, int *positions, float *values, int *first_active_start) {
    (void)cols;
    trip_key *keys = (trip_key *)malloc((size_t)(n_triplets > 0 ? n_triplets : 1) * sizeof(trip_key));
    trip_key *tmp = (trip_key *)malloc((size_t)(n_triplets > 0 ? n_triplets : 1) * sizeof(trip_key));
    for (int64_t i = 0; i < n_triplets; ++i) {
        keys[i].key = ((uint64_t)(uint32_t)trow[i] << 32) | (uint64_t)(uint32_t)tcol[i]; /* H:611 */
        keys[i].seq = i;
    }
    trip_merge_sort(keys, tmp, n_triplets);
    free(tmp);

    /* H:1615-1617 count per row (over distinct keys), H:1619-1628 prefix sum + firstActiveStart */
    for (int i = 0; i <= rows; ++i) start[i] = 0;
    int64_t nnz = 0;
    for (int64_t i = 0; i < n_triplets;) {
        int64_t j = i;
        float v = tval[keys[i].seq];                    /* H:614 first insert */
        for (j = i + 1; j < n_triplets && keys[j].key == keys[i].key; ++j) v += tval[keys[j].seq]; /* H:616 */
        int row = (int)(keys[i].key >> 32);             /* H:402-404 */
        positions[nnz] = (int)(keys[i].key & 0xFFFFFFFFu); /* H:406-408, H:1636 */
        values[nnz] = v;                                /* H:1637 */
        start[row + 1]++;
        nnz++;
        i = j;
    }
    int fas = -1;
    for (int i = 0; i < rows; ++i) {
        start[i + 1] += start[i];
        if (fas == -1 && start[i + 1] != 0) fas = i;    /* H:1622-1624 */
    }
    if (fas == -1) fas = rows;                          /* H:1626-1628 */
    if (first_active_start) *first_active_start = fas;
    free(keys);
    return (int)nnz;
}

/* ------------------------------------------------------------------------------------------------
 * CSRMatrix::rMultOp, H:1458-1499 (+ wrappers H:1501-1515)
 * ---------------------------------------------------------------------------------------------- */
void smm_oracle_spmv(int rows, const int *start, const int *positions, const float *values,
                     int op, const float *lhs, const float *mult, float *out) {
#pragma omp parallel for schedule(static)
    for (int row = 0; row < rows; ++row) {
        float dot = 0.0f;                               /* H:1484; empty row -> op(lhs, 0) H:1479-1483 */
        for (int k = start[row]; k < start[row + 1]; ++k)
            dot = smm_fma(values[k], mult[positions[k]], dot); /* H:1485-1489 */
        float l = (op == 0) ? 0.0f : lhs[row];
        out[row] = (op == 0) ? dot : (op == 1 ? l + dot : l - dot); /* H:1284-1286, H:1509, H:1514 */
    }
}

/* ------------------------------------------------------------------------------------------------
 * Vector::operator*, H:305-328
 *   serial: left-to-right.
 *   SMM_MULTITHREADING: tbb::parallel_deterministic_reduce over blocked_range<int>(0,size,8192) with
 *   identity 0.0f and std::plus: the range is halved at begin+(end-begin)/2 while size > grain
 *   (blocked_range::is_divisible), every leaf is summed sequentially from the identity, and joins are
 *   left + right.  The result therefore does not depend on the number of threads.
 * ---------------------------------------------------------------------------------------------- */
#define SMM_DOT_GRAIN 8192

static float dot_leaf(const float *a, const float *b, int begin, int end) {
    float cur = 0.0f;
    for (int j = begin; j < end; ++j) cur += a[j] * b[j]; /* H:314-316 */
    return cur;
}

static float dot_tree(const float *a, const float *b, int begin, int end, int depth) {
    if (end - begin > SMM_DOT_GRAIN) {
        int mid = begin + (end - begin) / 2;
        float l, r;
        if (depth < 6) {
#pragma omp task shared(l) if (end - begin > (1 << 16))
            l = dot_tree(a, b, begin, mid, depth + 1);
            r = dot_tree(a, b, mid, end, depth + 1);
#pragma omp taskwait
        } else {
            l = dot_tree(a, b, begin, mid, depth + 1);
            r = dot_tree(a, b, mid, end, depth + 1);
        }
        return l + r;                                   /* std::plus join, H:319 */
    }
    return dot_leaf(a, b, begin, end);
}

float smm_oracle_dot(int n, const float *a, const float *b, int dot_mode) {
    if (dot_mode == SMM_ORACLE_DOT_SERIAL) return dot_leaf(a, b, 0, n); /* H:322-326 */
    float res = 0.0f;
    if (n <= (1 << 16)) return dot_tree(a, b, 0, n, 99);
#pragma omp parallel
#pragma omp single
    res = dot_tree(a, b, 0, n, 0);
    return res;
}

/* ------------------------------------------------------------------------------------------------
 * SGSPreconditioner::apply, H:1658-1713
 * ---------------------------------------------------------------------------------------------- */
int smm_oracle_sgs_apply(int rows, const int *start, const int *positions, const float *values,
                         int first_active_start, const float *rhs, float *x) {
    if (first_active_start != 0) return 1;              /* H:1668-1670 */
    for (int row = 0; row < rows; ++row) {              /* forward, H:1673-1695 */
        int k = start[row];
        if (start[row + 1] - k == 0) return 1;          /* H:1678-1680 */
        int col = positions[k];
        float value = values[k];
        float lhs = rhs[row];
        while (col < row) {
            lhs = smm_fma(-value, x[col], lhs);         /* H:1685 */
            ++k;
            /* The reference reads positions[k] unconditionally (H:1687); when a row has no entry at or
             * right of the diagonal this walks into the next row (or one past the arrays on the last
             * row).  Stop at the end of the arrays instead of reading out of bounds; the result is the
             * same error code. */
            if (k >= start[rows]) return 1;
            col = positions[k];
            value = values[k];
        }
        if (col != row || fabsf(value) < 1e-5) return 1; /* H:1691-1693 (double literal: float promoted) */
        x[row] = lhs / value;                           /* H:1694 */
    }
    for (int row = rows - 1; row >= 0; --row) {         /* backward, H:1698-1711 */
        int k = start[row + 1] - 1;
        int col = positions[k];
        float value = values[k];
        float lhs = 0.0f;
        while (col > row) {
            lhs = smm_fma(value, x[col], lhs);          /* H:1704 */
            --k;
            col = positions[k];
            value = values[k];
        }
        x[row] = x[row] - lhs / value;                  /* H:1710 */
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * IC0Preconditioner::factorize (H:1839-1928) and ::apply (H:1802-1837)
 * ---------------------------------------------------------------------------------------------- */
int smm_oracle_ic0_factorize(int rows, const int *start, const int *positions, const float *values, float *ic0) {
    int *next_free = (int *)calloc((size_t)(rows > 0 ? rows : 1), sizeof(int));
    int *used = (int *)malloc((size_t)(rows > 0 ? rows : 1) * sizeof(int));
    for (int i = 0; i < rows; ++i) used[i] = -1;
    int rc = 0;
    for (int i = 0; i < rows && rc == 0; ++i) {
        for (int j = start[i]; j < start[i + 1]; ++j) used[positions[j]] = j;   /* H:1860-1863 */
        float diag = 0.0f;
        int ci = start[i];
        int column = positions[ci];
        while (column < i) {                              /* H:1868-1872 */
            diag += ic0[ci] * ic0[ci];
            ci++;
            column = positions[ci];
        }
        if (column != i) { rc = 1; break; }               /* H:1873-1876 */
        const int diag_pos = start[i] + next_free[i];     /* H:1877 */
        diag = sqrtf(values[ci] - diag);                  /* H:1879 */
        ic0[diag_pos] = diag;
        next_free[i]++;
        const float diag_inv = 1.0f / diag;               /* H:1883 */
        for (int j = i + 1; j < rows; ++j) {              /* H:1888 (quadratic, as in the reference) */
            const int row_start = start[j];
            const int value_index = row_start + next_free[j];
            if (value_index >= start[rows] || positions[value_index] != i) continue;  /* H:1893 */
            float sum = 0.0f;
            const int row_end = start[j + 1];
            int k = row_start;
            int col = positions[k];
            while (k < row_end && col < i) {              /* H:1900-1907 */
                const int iv = used[col];
                if (iv != -1) sum += ic0[iv] * ic0[k];
                k++;
                col = positions[k];
            }
            sum = (values[k] - sum) * diag_inv;           /* H:1914 */
            ic0[k] = sum;
            ic0[start[i] + next_free[i]] = sum;           /* H:1916-1917 */
            next_free[i]++;
            next_free[j]++;
        }
        for (int j = start[i]; j < start[i + 1]; ++j) used[positions[j]] = -1;  /* H:1922-1925 */
    }
    free(next_free);
    free(used);
    return rc;
}

int smm_oracle_ic0_apply(int rows, const int *start, const int *positions, const float *ic0,
                         const float *rhs, float *x) {
    for (int row = 0; row < rows; ++row) {                /* H:1806-1819 */
        float sum = rhs[row];
        int j = start[row];
        const int row_end = start[row + 1];
        int col = positions[j];
        while (col < row && j < row_end) {
            sum -= ic0[j] * x[col];
            j++;
            col = positions[j];
        }
        x[row] = sum / ic0[j];
    }
    for (int row = rows - 1; row >= 0; --row) {           /* H:1822-1835 */
        float sum = x[row];
        const int row_start = start[row];
        int j = start[row + 1] - 1;
        int col = positions[j];
        while (col > row && j >= row_start) {
            sum -= ic0[j] * x[col];
            j--;
            col = positions[j];
        }
        x[row] = sum / ic0[j];
    }
    return 0;
}

/* EXTENSION (parity UNPINNED by the reference): zero-fill incomplete LU as ILU0Preconditioner::factorize describes it
 * (H:1723-1790): row-wise IKJ on A's pattern, unit L with the diagonal implied, U's diagonal stored, multiplier
 * alphaIK = ilu0Val[kPos] * diagonalElementsInv[k] (H:1762), update ilu0Val[..] -= alphaIK * betaKJ (H:1766-1768).
 * The reference's code cannot run to completion (inverted guards at H:1744 / H:1775, inner loop bound `col > 0`
 * at H:1764 instead of `col > k`), and it has no apply(); this restates the algorithm those lines describe.
 * Returns 0, 1 (first_active_start != 0 / empty row / missing diagonal), 2 (pivot not > 1e-6 in magnitude). */
int smm_oracle_ilu0_factorize(int rows, const int *start, const int *positions, const float *values,
                              int first_active_start, float *ilu0) {
    if (rows == 0) return 0;
    if (first_active_start != 0) return 1;                 /* H:1734-1737 */
    const int nnz = start[rows];
    memcpy(ilu0, values, sizeof(float) * (size_t)nnz);     /* H:1732 */
    int *column_index = (int *)malloc(sizeof(int) * (size_t)rows);
    float *dinv = (float *)malloc(sizeof(float) * (size_t)rows);
    for (int i = 0; i < rows; ++i) column_index[i] = -1;
    int rc = 0;
    for (int row = 0; row < rows && rc == 0; ++row) {
        const int rs = start[row], re = start[row + 1];
        for (int i = rs; i < re; ++i) column_index[positions[i]] = i;
        int kp = rs;
        for (; kp < re && positions[kp] < row; ++kp) {
            const int k = positions[kp];
            const float alpha = ilu0[kp] * dinv[k];
            ilu0[kp] = alpha;
            for (int cp = start[k + 1] - 1; cp >= start[k] && positions[cp] > k; --cp) {
                const int ci = column_index[positions[cp]];
                if (ci != -1) ilu0[ci] -= alpha * ilu0[cp];
            }
        }
        for (int i = rs; i < re; ++i) column_index[positions[i]] = -1;
        if (kp >= re || positions[kp] != row) { rc = 1; break; }
        if (!(fabsf(ilu0[kp]) > 1e-6f)) { rc = 2; break; }
        dinv[row] = 1.0f / ilu0[kp];
    }
    free(column_index);
    free(dinv);
    return rc;
}

/* L y = rhs (unit diagonal, columns ascending), U x = y (columns descending, one division by u_ii): the same loop
 * shape as IC0Preconditioner::apply (H:1802-1837), which is the only factor-based apply the reference defines. */
int smm_oracle_ilu0_apply(int rows, const int *start, const int *positions, const float *ilu0,
                          const float *rhs, float *x) {
    for (int row = 0; row < rows; ++row) {
        float sum = rhs[row];
        for (int j = start[row]; j < start[row + 1] && positions[j] < row; ++j) sum -= ilu0[j] * x[positions[j]];
        x[row] = sum;
    }
    for (int row = rows - 1; row >= 0; --row) {
        float sum = x[row];
        int j = start[row + 1] - 1;
        for (; j >= start[row] && positions[j] > row; --j) sum -= ilu0[j] * x[positions[j]];
        x[row] = sum / ilu0[j];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Solvers
 * ---------------------------------------------------------------------------------------------- */
static float *vec_alloc(int n) { return (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float)); }
static float *vec_zero(int n) { return (float *)calloc((size_t)(n > 0 ? n : 1), sizeof(float)); }

static void push_history(float *history, int cap, int it, float v) {
    if (history && it < cap) history[it] = v;
}

/* ConjugateGradient, H:2316-2398 */
void smm_oracle_cg(int rows, const int *start, const int *positions, const float *values,
                   const float *b, const float *x0, float *x, int max_iterations, float eps, int mt,
                   smm_oracle_info *info, float *history, int history_cap) {
    const int dm = mt ? SMM_ORACLE_DOT_TBB8192 : SMM_ORACLE_DOT_SERIAL;
    const float eps2 = eps * eps;                          /* H:2335 */
    float *r = vec_zero(rows), *p = vec_alloc(rows), *Ap = vec_zero(rows);
    smm_oracle_spmv(rows, start, positions, values, 2, b, x0, r);   /* H:2337 */
    memcpy(p, r, (size_t)rows * sizeof(float));            /* H:2340 */
    float rr = smm_oracle_dot(rows, r, r, dm);             /* H:2341 */
    info->iterations = 0;
    info->precond_error = 0;
    info->residual = rr;
    if (eps2 > rr) { info->status = SMM_ORACLE_SUCCESS; goto done; }  /* H:2342-2344 */
    if (max_iterations == -1) max_iterations = rows;       /* H:2345-2347 */
    {
        const float *cur_x = x0;                           /* H:2351 */
        for (int i = 0; i < max_iterations; ++i) {
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, Ap); /* H:2353 */
            const float pAp = smm_oracle_dot(rows, Ap, p, dm);   /* H:2354 */
            const float alpha = rr / pAp;                  /* H:2358 */
            float new_rr = 0.0f;
            if (mt) {                                      /* H:2363-2369 */
#pragma omp parallel for schedule(static)
                for (int j = 0; j < rows; ++j) {
                    x[j] = smm_fma(alpha, p[j], cur_x[j]);
                    r[j] = smm_fma(-alpha, Ap[j], r[j]);
                }
                new_rr = smm_oracle_dot(rows, r, r, dm);
            } else {                                       /* H:2371-2375 */
                for (int j = 0; j < rows; ++j) {
                    x[j] = smm_fma(alpha, p[j], cur_x[j]);
                    r[j] = smm_fma(-alpha, Ap[j], r[j]);
                    new_rr += r[j] * r[j];
                }
            }
            info->iterations = i + 1;
            info->residual = new_rr;
            push_history(history, history_cap, i, new_rr);
            if (eps2 > new_rr) { info->status = SMM_ORACLE_SUCCESS; goto done; } /* H:2377-2379 */
            const float beta = new_rr / rr;                /* H:2381 */
            rr = new_rr;
#pragma omp parallel for schedule(static)
            for (int j = 0; j < rows; ++j) p[j] = smm_fma(beta, p[j], r[j]);   /* H:2385-2393 */
            cur_x = x;                                     /* H:2395 */
        }
    }
    info->status = SMM_ORACLE_MAX_ITERATIONS_REACHED;      /* H:2397 */
done:
    free(r); free(p); free(Ap);
}

static int clamp_max_iterations(int max_iterations, int rows) {
    /* H:2030-2033, H:2111-2114, H:2200-2203 */
    if (max_iterations > rows) max_iterations = rows;
    if (max_iterations == -1) max_iterations = rows;
    return max_iterations;
}

/* BiCGSymmetric, H:2021-2102 */
void smm_oracle_bicgsym(int rows, const int *start, const int *positions, const float *values,
                        const float *b, float *x, int max_iterations, float eps, int mt,
                        smm_oracle_info *info, float *history, int history_cap) {
    const int dm = mt ? SMM_ORACLE_DOT_TBB8192 : SMM_ORACLE_DOT_SERIAL;
    max_iterations = clamp_max_iterations(max_iterations, rows);
    float *r = vec_alloc(rows), *p = vec_alloc(rows), *ap = vec_alloc(rows);
    smm_oracle_spmv(rows, start, positions, values, 2, b, x, r);  /* H:2036 */
    memcpy(p, r, (size_t)rows * sizeof(float));
    float r2 = smm_oracle_dot(rows, r, r, dm);             /* H:2043 */
    int iterations = 0;
    const float eps2 = eps * eps;
    info->precond_error = 0;
    do {
        smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, ap);   /* H:2048 */
        const float denom = smm_oracle_dot(rows, ap, p, dm);   /* H:2049 */
        if (eps > fabsf(denom) && r2 > 1) {                /* H:2056-2058 */
            info->status = SMM_ORACLE_DIVERGED; info->iterations = iterations; info->residual = r2;
            goto done;
        }
        const float alpha = r2 / denom;                    /* H:2059 */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < rows; ++j) {                   /* H:2061-2071: plain `*` then `+=`, no _smm_fma */
            x[j] += alpha * p[j];
            r[j] -= alpha * ap[j];
        }
        const float new_r2 = smm_oracle_dot(rows, r, r, dm);   /* H:2075 */
        if (new_r2 > 1 && r2 < eps) {                      /* H:2079-2081 */
            info->status = SMM_ORACLE_DIVERGED; info->iterations = iterations; info->residual = new_r2;
            goto done;
        }
        const float beta = new_r2 / r2;                    /* H:2082 */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < rows; ++j) p[j] = r[j] + beta * p[j];   /* H:2084-2092 */
        r2 = new_r2;
        push_history(history, history_cap, iterations, r2);
        iterations++;
    } while (r2 > eps2 && iterations < max_iterations);    /* H:2096 */
    info->iterations = iterations;
    info->residual = r2;
    info->status = (iterations > max_iterations) ? SMM_ORACLE_MAX_ITERATIONS_REACHED : SMM_ORACLE_SUCCESS;
done:
    free(r); free(p); free(ap);
}

/* ConjugateGradientSquared, H:2109-2178 (with residualSquared hoisted so that it compiles) */
void smm_oracle_cgs(int rows, const int *start, const int *positions, const float *values,
                    const float *b, float *x, int max_iterations, float eps, int mt,
                    smm_oracle_info *info, float *history, int history_cap) {
    const int dm = mt ? SMM_ORACLE_DOT_TBB8192 : SMM_ORACLE_DOT_SERIAL;
    max_iterations = clamp_max_iterations(max_iterations, rows);
    float *r = vec_alloc(rows), *r0 = vec_alloc(rows), *p = vec_alloc(rows), *u = vec_alloc(rows);
    float *q = vec_alloc(rows), *auq = vec_alloc(rows), *ap = vec_alloc(rows);
    smm_oracle_spmv(rows, start, positions, values, 2, b, x, r);  /* H:2118 */
    memcpy(p, r, (size_t)rows * sizeof(float));
    memcpy(u, r, (size_t)rows * sizeof(float));
    memcpy(r0, r, (size_t)rows * sizeof(float));
    float rr0 = smm_oracle_dot(rows, r, r0, dm);           /* H:2128 */
    int iterations = 0;
    const float eps2 = eps * eps;
    float res2 = 0.0f;
    info->precond_error = 0;
    do {
        smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, ap);   /* H:2132 */
        const float denom = smm_oracle_dot(rows, ap, r0, dm);  /* H:2133 */
        const float alpha = rr0 / denom;                   /* H:2135 */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < rows; ++j) {                   /* H:2137-2149 */
            q[j] = smm_fma(-alpha, ap[j], u[j]);
            auq[j] = alpha * (u[j] + q[j]);
            x[j] = x[j] + auq[j];
        }
        smm_oracle_spmv(rows, start, positions, values, 2, r, auq, r);     /* H:2151 in place */
        const float new_rr0 = smm_oracle_dot(rows, r, r0, dm); /* H:2152 */
        const float beta = new_rr0 / rr0;                  /* H:2154 */
#pragma omp parallel for schedule(static)
        for (int j = 0; j < rows; ++j) {                   /* H:2157-2167 */
            u[j] = smm_fma(beta, q[j], r[j]);
            p[j] = smm_fma(beta, smm_fma(beta, p[j], q[j]), u[j]);
        }
        rr0 = new_rr0;
        res2 = smm_oracle_dot(rows, r, r, dm);             /* H:2171 */
        push_history(history, history_cap, iterations, res2);
        iterations++;
    } while (res2 > eps2 && iterations < max_iterations);  /* H:2172 */
    info->iterations = iterations;
    info->residual = res2;
    info->status = (iterations > max_iterations) ? SMM_ORACLE_MAX_ITERATIONS_REACHED : SMM_ORACLE_SUCCESS;
    free(r); free(r0); free(p); free(u); free(q); free(auq); free(ap);
}

/* BiCGStab, H:2191-2283 (+ wrapper H:2294-2303) */
/* EXTENSION (not in the reference): diagonal preconditioner, x_i = rhs_i / a_ii; 1 when a row has no diagonal entry */
int smm_oracle_jacobi_apply(int rows, const int *start, const int *positions, const float *values,
                            const float *rhs, float *x) {
    for (int row = 0; row < rows; ++row) {
        int j = start[row];
        while (j < start[row + 1] && positions[j] < row) ++j;
        if (j >= start[row + 1] || positions[j] != row) return 1;
        x[row] = rhs[row] / values[j];
    }
    return 0;
}

static int apply_precond(int kind, const float *factor, int rows, const int *start, const int *positions, const float *values,
                         int first_active_start, const float *rhs, float *x) {
    if (kind == 1) return smm_oracle_sgs_apply(rows, start, positions, values, first_active_start, rhs, x);
    if (kind == 2) return smm_oracle_ilu0_apply(rows, start, positions, factor, rhs, x);
    if (kind == 4) return smm_oracle_jacobi_apply(rows, start, positions, values, rhs, x);
    return smm_oracle_ic0_apply(rows, start, positions, factor, rhs, x);
}

void smm_oracle_bicgstab(int rows, const int *start, const int *positions, const float *values,
                         int first_active_start, int precond,
                         const float *b, float *x, int max_iterations, float eps, int mt,
                         smm_oracle_info *info, float *history, int history_cap) {
    smm_oracle_bicgstab_pc(rows, start, positions, values, first_active_start, precond, NULL, b, x, max_iterations, eps, mt,
                           info, history, history_cap);
}

/* the template instantiated with any preconditioner object (H:2191-2199): precond 2 = ILU(0), 3 = IC(0), factor = its values */
void smm_oracle_bicgstab_pc(int rows, const int *start, const int *positions, const float *values,
                            int first_active_start, int precond, const float *factor,
                            const float *b, float *x, int max_iterations, float eps, int mt,
                            smm_oracle_info *info, float *history, int history_cap) {
    const int dm = mt ? SMM_ORACLE_DOT_TBB8192 : SMM_ORACLE_DOT_SERIAL;
    max_iterations = clamp_max_iterations(max_iterations, rows);
    float *scratch = precond ? vec_alloc(rows) : NULL;     /* H:2208-2212 */
    float *r = vec_alloc(rows), *r0 = vec_alloc(rows), *p = vec_alloc(rows);
    float *ap = vec_alloc(rows), *s = vec_alloc(rows), *as = vec_alloc(rows);
    int perr = 0;
    smm_oracle_spmv(rows, start, positions, values, 2, b, x, r);  /* H:2215 */
    if (precond) perr |= apply_precond(precond, factor, rows, start, positions, values, first_active_start, r, scratch); /* H:2218 */
    for (int i = 0; i < rows; ++i) {                       /* H:2221-2227 */
        if (precond) r[i] = scratch[i];
        r0[i] = r[i];
        p[i] = r[i];
    }
    float res = 0.0f;
    int iterations = 0;
    float rr0 = smm_oracle_dot(rows, r, r0, dm);           /* H:2231 */
    do {
        if (precond) {                                     /* H:2233-2241 */
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, scratch);
            perr |= apply_precond(precond, factor, rows, start, positions, values, first_active_start, scratch, ap);
        } else {
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, ap);
        }
        float denom = smm_oracle_dot(rows, ap, r0, dm);    /* H:2243 */
        const float alpha = rr0 / denom;                   /* H:2244 */
        for (int i = 0; i < rows; ++i) s[i] = smm_fma(-alpha, ap[i], r[i]);   /* H:2245-2247 */
        if (precond) {                                     /* H:2249-2257 */
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, s, scratch);
            perr |= apply_precond(precond, factor, rows, start, positions, values, first_active_start, scratch, as);
        } else {
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, s, as);
        }
        denom = smm_oracle_dot(rows, as, as, dm);          /* H:2259 */
        const float omega = smm_oracle_dot(rows, as, s, dm) / denom;   /* H:2261 */
        res = 0.0f;
        for (int i = 0; i < rows; ++i) {                   /* H:2263-2267, serial in both builds */
            x[i] = smm_fma(alpha, p[i], smm_fma(omega, s[i], x[i]));
            r[i] = smm_fma(-omega, as[i], s[i]);
            res += r[i] * r[i];
        }
        res = sqrtf(res);                                  /* H:2268 */
        const float new_rr0 = smm_oracle_dot(rows, r, r0, dm);  /* H:2269 */
        const float beta = (new_rr0 * alpha) / (rr0 * omega);   /* H:2271 */
        for (int i = 0; i < rows; ++i) p[i] = smm_fma(beta, smm_fma(-omega, ap[i], p[i]), r[i]); /* H:2272-2274 */
        rr0 = new_rr0;
        push_history(history, history_cap, iterations, res);
        iterations++;
    } while (res > eps && iterations < max_iterations);    /* H:2277 */
    info->iterations = iterations;
    info->residual = res;
    info->precond_error = perr;
    info->status = (iterations > max_iterations) ? SMM_ORACLE_MAX_ITERATIONS_REACHED : SMM_ORACLE_SUCCESS;
    free(scratch); free(r); free(r0); free(p); free(ap); free(s); free(as);
}

/* ConjugateGradient with IC0, H:2414-2505 */
void smm_oracle_cg_ic0(int rows, const int *start, const int *positions, const float *values,
                       const float *ic0, const float *b, const float *x0, float *x,
                       int max_iterations, float eps, int mt,
                       smm_oracle_info *info, float *history, int history_cap) {
    const int dm = mt ? SMM_ORACLE_DOT_TBB8192 : SMM_ORACLE_DOT_SERIAL;
    const float eps2 = eps * eps;
    float *r = vec_zero(rows), *z = vec_zero(rows), *p = vec_zero(rows), *Ap = vec_zero(rows);
    smm_oracle_spmv(rows, start, positions, values, 2, b, x0, r);   /* H:2440 */
    smm_oracle_ic0_apply(rows, start, positions, ic0, r, z);        /* H:2441 */
    float rz = 0.0f, rr = 0.0f;
    for (int i = 0; i < rows; ++i) {                       /* H:2444-2448 */
        rz += r[i] * z[i];
        rr += r[i] * r[i];
        p[i] = z[i];
    }
    info->iterations = 0;
    info->precond_error = 0;
    info->residual = rr;
    if (eps2 > rr) { info->status = SMM_ORACLE_SUCCESS; goto done; }
    if (max_iterations == -1) max_iterations = rows;
    {
        const float *cur_x = x0;
        for (int i = 0; i < max_iterations; ++i) {
            smm_oracle_spmv(rows, start, positions, values, 0, NULL, p, Ap);   /* H:2461 */
            const float pAp = smm_oracle_dot(rows, Ap, p, dm);
            const float alpha = rz / pAp;                  /* H:2466 */
#pragma omp parallel for schedule(static)
            for (int j = 0; j < rows; ++j) {               /* H:2470-2480 */
                x[j] = smm_fma(alpha, p[j], cur_x[j]);
                r[j] = smm_fma(-alpha, Ap[j], r[j]);
            }
            smm_oracle_ic0_apply(rows, start, positions, ic0, r, z);   /* H:2482 */
            const float new_rz = smm_oracle_dot(rows, r, z, dm);       /* H:2483 */
            rr = smm_oracle_dot(rows, r, r, dm);           /* H:2484 */
            info->iterations = i + 1;
            info->residual = rr;
            push_history(history, history_cap, i, rr);
            if (eps2 > rr) { info->status = SMM_ORACLE_SUCCESS; goto done; }
            const float beta = new_rz / rz;                /* H:2488 */
#pragma omp parallel for schedule(static)
            for (int j = 0; j < rows; ++j) p[j] = smm_fma(beta, p[j], z[j]);   /* H:2490-2499 */
            rz = new_rz;
            cur_x = x;
        }
    }
    info->status = SMM_ORACLE_MAX_ITERATIONS_REACHED;
done:
    free(r); free(z); free(p); free(Ap);
}

/* ------------------------------------------------------------------------------------------------
 * loadMatrixMarketMatrix, H:2531-2609.  MatrixLoadStatus values, H:2507-2522:
 *   0 SUCCESS, 1 FAILED_TO_OPEN_FILE, 2 UNKNOWN_FORMAT, 3 FAILED_TO_PARSE_FILE, 4 MISSING_BANNER,
 *   5 UNSUPPORTED_TYPE, 6 UNSUPPORTED_FORMAT, 7 UNSUPPORTED_EL_TYPE, 8 UNSUPPORTED_STRUCTURE
 * ---------------------------------------------------------------------------------------------- */
static int read_token(FILE *f, char *buf, int cap) {
    int c;
    do { c = fgetc(f); } while (c != EOF && isspace(c));    /* operator>> skips leading whitespace */
    if (c == EOF) return 0;
    int n = 0;
    while (c != EOF && !isspace(c)) {
        if (n < cap - 1) buf[n++] = (char)c;
        c = fgetc(f);
    }
    if (c != EOF) ungetc(c, f);
    buf[n] = 0;
    return 1;
}

static void lower(char *s) { for (; *s; ++s) *s = (char)tolower((unsigned char)*s); }

static int peekc(FILE *f) { int c = fgetc(f); if (c != EOF) ungetc(c, f); return c; }

static void ignore_line(FILE *f) { int c; do { c = fgetc(f); } while (c != EOF && c != '\n'); }

int smm_oracle_load_mtx(const char *path, int *rows, int *cols, int64_t *n_triplets,
                        int **trow, int **tcol, float **tval) {
    *trow = *tcol = NULL; *tval = NULL; *n_triplets = 0; *rows = *cols = 0;
    FILE *f = fopen(path, "r");
    if (!f) return 1;                                      /* H:2534-2536 */
    char tok[256];
    int rc = 0;
    if (!read_token(f, tok, sizeof tok) || strcmp(tok, "%%MatrixMarket") != 0) { rc = 4; goto out; } /* H:2546-2549 */
    if (!read_token(f, tok, sizeof tok)) tok[0] = 0;
    lower(tok);
    if (strcmp(tok, "matrix") != 0) { rc = 5; goto out; }  /* H:2551-2555 */
    if (!read_token(f, tok, sizeof tok)) tok[0] = 0;
    lower(tok);
    if (strcmp(tok, "coordinate") != 0) { rc = 6; goto out; } /* H:2557-2561 */
    if (!read_token(f, tok, sizeof tok)) tok[0] = 0;
    lower(tok);
    if (strcmp(tok, "real") != 0 && strcmp(tok, "integer") != 0) { rc = 7; goto out; } /* H:2563-2567 */
    if (!read_token(f, tok, sizeof tok)) tok[0] = 0;
    lower(tok);
    if (strcmp(tok, "symmetric") != 0) { rc = 8; goto out; } /* H:2569-2573 */
    for (int c = peekc(f); c == '%' || (c != EOF && isspace(c)); c = peekc(f)) ignore_line(f); /* H:2576-2578 */
    int nnz;
    if (fscanf(f, "%d %d %d", rows, cols, &nnz) != 3) { rc = 3; goto out; }  /* H:2580-2584 */
    {
        int64_t cap = 2 * (int64_t)(nnz > 0 ? nnz : 1), n = 0;
        *trow = (int *)malloc((size_t)cap * sizeof(int));
        *tcol = (int *)malloc((size_t)cap * sizeof(int));
        *tval = (float *)malloc((size_t)cap * sizeof(float));
        /* H:2588 `while (!file.eof())`: eofbit is only set once a read hit end-of-file, so after the
         * size line the loop body always runs at least once. */
        int at_eof = 0;
        while (!at_eof) {
            int r, c;
            char num[128];
            /* operator>>(float&) parses the decimal text directly to float, correctly rounded */
            if (fscanf(f, "%d %d", &r, &c) != 2 || !read_token(f, num, sizeof num)) { rc = 3; goto out; } /* H:2591-2594 */
            char *endp;
            float v = strtof(num, &endp);
            if (endp == num) { rc = 3; goto out; }
            r -= 1; c -= 1;                                /* H:2596 */
            if (n + 2 > cap) {
                cap *= 2;
                *trow = (int *)realloc(*trow, (size_t)cap * sizeof(int));
                *tcol = (int *)realloc(*tcol, (size_t)cap * sizeof(int));
                *tval = (float *)realloc(*tval, (size_t)cap * sizeof(float));
            }
            (*trow)[n] = r; (*tcol)[n] = c; (*tval)[n] = v; n++;          /* H:2598 */
            if (r != c) { (*trow)[n] = c; (*tcol)[n] = r; (*tval)[n] = v; n++; } /* H:2599-2601 */
            /* H:2603-2605; a value token that ended exactly at end-of-file sets eofbit in the reference's
             * stream (num_get reads until a non-numeric char), which ends the loop. */
            int pc = peekc(f);
            if (pc == EOF) at_eof = 1;
            while (pc != EOF && isspace(pc)) { ignore_line(f); pc = peekc(f); if (pc == EOF) at_eof = 1; }
        }
        *n_triplets = n;
    }
out:
    fclose(f);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * benchmark inputs (not part of the reference: its only ingest path is the std::map TripletMatrix)
 * ---------------------------------------------------------------------------------------------- */
int64_t smm_oracle_stencil_nnz(int nx, int ny, int nz, int use_z) {
    const int64_t rows = (int64_t)nx * ny * nz;
    const int full = use_z ? 7 : 5;
    return rows * full - 2ll * ny * nz - 2ll * nx * nz - (use_z ? 2ll * nx * ny : 0);
}

static int64_t stencil_start(int64_t r, int64_t nx, int64_t ny, int64_t nz, int use_z) {
    const int64_t plane = nx * ny, k = r / plane, rem = r % plane, lines = r / nx, i = r % nx;
    const int full = use_z ? 7 : 5;
    int64_t miss = lines + (i > 0 ? 1 : 0) + lines;
    miss += k * nx + (rem < nx ? rem : nx);
    { const int64_t last = rem - (ny - 1) * nx; miss += k * nx + (last > 0 ? last : 0); }
    if (use_z) {
        miss += (r < plane ? r : plane);
        { const int64_t last = r - (nz - 1) * plane; miss += (last > 0 ? last : 0); }
    }
    return r * full - miss;
}

void smm_oracle_gen_stencil(int nx, int ny, int nz, int use_z, float lo, float diag, float hi,
                            int *start, int *positions, float *values) {
    const int64_t rows = (int64_t)nx * ny * nz, plane = (int64_t)nx * ny;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r <= rows; ++r) {
        int64_t o = stencil_start(r, nx, ny, nz, use_z);
        start[r] = (int)o;
        if (r == rows) continue;
        const int64_t k = r / plane, j = (r % plane) / nx, i = r % nx;
        if (use_z && k > 0) { positions[o] = (int)(r - plane); values[o] = lo; ++o; }
        if (j > 0) { positions[o] = (int)(r - nx); values[o] = lo; ++o; }
        if (i > 0) { positions[o] = (int)(r - 1); values[o] = lo; ++o; }
        positions[o] = (int)r; values[o] = diag; ++o;
        if (i < nx - 1) { positions[o] = (int)(r + 1); values[o] = hi; ++o; }
        if (j < ny - 1) { positions[o] = (int)(r + nx); values[o] = hi; ++o; }
        if (use_z && k < nz - 1) { positions[o] = (int)(r + plane); values[o] = hi; ++o; }
    }
}

void smm_oracle_gen_xstar(int64_t n, uint64_t seed, float *x) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint64_t z = seed + ((uint64_t)i + 1ull) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        x[i] = (float)(z >> 40) / 16777216.0f;
    }
}
