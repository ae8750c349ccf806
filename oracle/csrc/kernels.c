/* TEST / BASELINE INFRASTRUCTURE ONLY — multi-threaded CPU kernels for the oracle (oracle/krylov_oracle.py).
 *
 * The numpy oracle is single-threaded wherever scipy / numpy are (CSR matvec, the Python loops over the
 * basis).  KrylovKit itself runs its `Array` fast path multi-threaded (src/orthonormal.jl:151-196 unproject,
 * :236-275 rank1update, :322-354 basistransform, :96-110 project over Threads), so a fair CPU baseline needs
 * threads too.  These kernels restate those loops with OpenMP; oracle/native.py plugs them into the oracle's
 * primitives, the drivers (restart logic, convergence tests) stay the oracle's.  Never linked into the
 * product; only tests/, bench.py's CPU legs and __graft_entry__.build() touch this file.
 *
 * Where a loop order matters for rounding it follows the reference: unproject / basistransform accumulate
 * over the basis index j in increasing order per element, exactly like unproject_linear_kernel!
 * (orthonormal.jl:176-196) and basistransform_linear_multithreaded! (:336-346).  Reductions (dot, project)
 * use fixed-size row blocks summed in block order, so results do not depend on the thread count. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define RED_BLOCK 16384   /* rows per partial sum */

void nat_set_num_threads(int t) {
#ifdef _OPENMP
    if (t >= 1) omp_set_num_threads(t);
#else
    (void)t;
#endif
}

int nat_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* apply(A::CSR, x) — rows in parallel, each row summed in storage order */
void nat_spmv_csr_f64(int64_t n_rows, const int64_t* rowptr, const int32_t* colidx, const double* vals,
                      const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_rows; ++i) {
        double acc = 0.0;
        for (int64_t p = rowptr[i]; p < rowptr[i + 1]; ++p) acc += vals[p] * x[colidx[p]];
        y[i] = acc;
    }
}

/* inner(x, y): block partials, then one ordered pass over the partials */
double nat_dot_f64(int64_t n, const double* x, const double* y) {
    const int64_t nb = (n + RED_BLOCK - 1) / RED_BLOCK;
    double* part = (double*)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1));
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nb; ++b) {
        const int64_t i0 = b * RED_BLOCK, i1 = (i0 + RED_BLOCK < n) ? i0 + RED_BLOCK : n;
        double acc = 0.0;
#pragma omp simd reduction(+ : acc)
        for (int64_t i = i0; i < i1; ++i) acc += x[i] * y[i];
        part[b] = acc;
    }
    double s = 0.0;
    for (int64_t b = 0; b < nb; ++b) s += part[b];
    free(part);
    return s;
}

/* add!!(y, x, a): y <- y + a x */
void nat_axpy_f64(int64_t n, double a, const double* restrict x, double* restrict y) {
#pragma omp parallel for simd schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] += a * x[i];
}

/* scale!!(y, x, a): y <- a x */
void nat_scale_f64(int64_t n, double a, const double* x, double* y) {  /* y may alias x */
#pragma omp parallel for simd schedule(static)
    for (int64_t i = 0; i < n; ++i) y[i] = a * x[i];
}

/* project!!(h, b, x, alpha, beta): h[j] = beta*h[j] + alpha*<b[j], x>  (orthonormal.jl:88-118).
 * Row blocks in parallel, every block produces k partials from ONE pass over its rows of x, so x is read
 * once per block instead of once per basis vector. */
void nat_project_f64(int64_t n, int32_t k, const double* const* b, const double* x, double alpha, double beta,
                     double* h) {
    const int64_t nb = (n + RED_BLOCK - 1) / RED_BLOCK;
    double* part = (double*)malloc(sizeof(double) * (size_t)(nb > 0 ? nb : 1) * (size_t)k);
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nb; ++blk) {
        const int64_t i0 = blk * RED_BLOCK, i1 = (i0 + RED_BLOCK < n) ? i0 + RED_BLOCK : n;
        for (int32_t j = 0; j < k; ++j) {
            const double* q = b[j];
            double acc = 0.0;
#pragma omp simd reduction(+ : acc)
            for (int64_t i = i0; i < i1; ++i) acc += q[i] * x[i];
            part[(size_t)blk * k + j] = acc;
        }
    }
    for (int32_t j = 0; j < k; ++j) {
        double s = 0.0;
        for (int64_t blk = 0; blk < nb; ++blk) s += part[(size_t)blk * k + j];
        h[j] = (beta == 0.0 ? 0.0 : beta * h[j]) + alpha * s;
    }
    free(part);
}

static int64_t prevpow2(int64_t v) {
    int64_t p = 1;
    while (2 * p <= v) p *= 2;
    return p;
}

/* unproject_linear_multithreaded! (orthonormal.jl:151-196): y = beta*y + alpha * sum_j b[j]*c[j], row blocks of
 * prevpow(2, 4096 / k) elements spread over the threads, j outer / i inner inside a block */
void nat_unproject_f64(int64_t n, int32_t k, const double* const* b, const double* c, double alpha, double beta,
                       double* y) {
    const int64_t bs = prevpow2(4096 / (k > 0 ? k : 1) > 0 ? 4096 / (k > 0 ? k : 1) : 1);
    const int64_t nb = (n + bs - 1) / bs;
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nb; ++blk) {
        const int64_t i0 = blk * bs, i1 = (i0 + bs < n) ? i0 + bs : n;
        if (beta == 0.0) {
            for (int64_t i = i0; i < i1; ++i) y[i] = 0.0;
        } else if (beta != 1.0) {
            for (int64_t i = i0; i < i1; ++i) y[i] *= beta;
        }
        for (int32_t j = 0; j < k; ++j) {
            const double cj = c[j] * alpha;
            const double* q = b[j];
#pragma omp simd
            for (int64_t i = i0; i < i1; ++i) y[i] += q[i] * cj;
        }
    }
}

/* basistransform_linear_multithreaded! (orthonormal.jl:322-354): out[j] = sum_k b[k]*U[k,j], j < keep; U is
 * column-major m x keep; the caller installs out[j] as the new b[j] */
void nat_basistransform_f64(int64_t n, int32_t m, int32_t keep, const double* const* b, const double* U,
                            double* const* out) {
    const int64_t bs = prevpow2(4096 / (m > 0 ? m : 1) > 0 ? 4096 / (m > 0 ? m : 1) : 1);
    const int64_t nb = (n + bs - 1) / bs;
#pragma omp parallel for schedule(static)
    for (int64_t blk = 0; blk < nb; ++blk) {
        const int64_t i0 = blk * bs, i1 = (i0 + bs < n) ? i0 + bs : n;
        for (int32_t j = 0; j < keep; ++j) {
            double* o = out[j];
            for (int64_t i = i0; i < i1; ++i) o[i] = 0.0;
            for (int32_t kk = 0; kk < m; ++kk) {
                const double u = U[(size_t)j * m + kk];
                const double* q = b[kk];
#pragma omp simd
                for (int64_t i = i0; i < i1; ++i) o[i] += q[i] * u;
            }
        }
    }
}
