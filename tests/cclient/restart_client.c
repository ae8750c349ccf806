/* Plain-C client of include/b200krylov.h, second program: one whole thick-restart cycle of
 * eigsolve(A, x0, howmany, :SR, Lanczos(krylovdim, maxiter = 2, tol = 0)) — src/eigsolve/lanczos.jl:33-116 —
 * driven the way the benchmark path drives it, with no Python in between:
 *
 *   initialize (lanczos.jl:180-222)                       b2k_vec_* / b2k_op_apply_dot
 *   expand! until K = krylovdim (:45-79)                  b2k_lanczos_expand_many   (device-chained steps)
 *   T = U D U'  (tridiageigh!, dense/linalg.jl:109)        cyclic Jacobi below (the host's job; Julia calls LAPACK)
 *   restore Lanczos form (:88-105)                        b2k_host_lanczos_restart
 *   basistransform!(B, U[:, 1:keep]) (:106)               b2k_basis_transform
 *   B[keep+1] = r/beta; shrink! (:107-112)                b2k_vec_scale + handle bookkeeping
 *   expand! again until K = krylovdim                     b2k_lanczos_expand_many
 *   T = U D U'                                            -> the `howmany` smallest Ritz values, printed
 *
 * tests/test_cclient.py compares them with the oracle's eigsolve on the same (A, x0).
 * Usage: restart_client nx ny krylovdim howmany */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "b200krylov.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int32_t st_ = (call);                                                        \
        if (st_ != B2K_OK) {                                                         \
            fprintf(stderr, "%s failed (%d): %s\n", #call, st_, b2k_last_error(ctx)); \
            return 2;                                                                \
        }                                                                            \
    } while (0)

/* eigen-decomposition of the symmetric tridiagonal (dv, ev), K <= 64: cyclic Jacobi on the dense matrix;
 * D ascending, U column-major K x K with T U = U D */
static void tridiag_eig(int K, const double* dv, const double* ev, double* D, double* U) {
    double* A = (double*)calloc((size_t)K * K, sizeof(double));
    int i, j, p, q, sweep;
    for (i = 0; i < K; ++i) {
        A[i * K + i] = dv[i];
        if (i + 1 < K) { A[i * K + i + 1] = ev[i]; A[(i + 1) * K + i] = ev[i]; }
    }
    memset(U, 0, sizeof(double) * (size_t)K * K);
    for (i = 0; i < K; ++i) U[i * K + i] = 1.0;
    for (sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (p = 0; p < K; ++p) for (q = p + 1; q < K; ++q) off += A[p * K + q] * A[p * K + q];
        if (off < 1e-300) break;
        for (p = 0; p < K; ++p)
            for (q = p + 1; q < K; ++q) {
                const double apq = A[p * K + q];
                double theta, t, c, s;
                if (fabs(apq) < 1e-300) continue;
                theta = (A[q * K + q] - A[p * K + p]) / (2.0 * apq);
                t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                c = 1.0 / sqrt(t * t + 1.0);
                s = t * c;
                for (i = 0; i < K; ++i) {          /* A <- A J (columns p, q) */
                    const double aip = A[i * K + p], aiq = A[i * K + q];
                    A[i * K + p] = c * aip - s * aiq;
                    A[i * K + q] = s * aip + c * aiq;
                }
                for (i = 0; i < K; ++i) {          /* A <- J' A (rows p, q) */
                    const double api = A[p * K + i], aqi = A[q * K + i];
                    A[p * K + i] = c * api - s * aqi;
                    A[q * K + i] = s * api + c * aqi;
                }
                for (i = 0; i < K; ++i) {          /* U <- U J; U stored column-major: U[col*K + row] */
                    const double uip = U[p * K + i], uiq = U[q * K + i];
                    U[p * K + i] = c * uip - s * uiq;
                    U[q * K + i] = s * uip + c * uiq;
                }
            }
    }
    for (i = 0; i < K; ++i) D[i] = A[i * K + i];
    for (i = 0; i < K; ++i)                        /* selection sort ascending, permuting the columns of U */
        for (j = i + 1; j < K; ++j)
            if (D[j] < D[i]) {
                double td = D[i];
                D[i] = D[j];
                D[j] = td;
                for (p = 0; p < K; ++p) {
                    const double tu = U[i * K + p];
                    U[i * K + p] = U[j * K + p];
                    U[j * K + p] = tu;
                }
            }
    free(A);
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s nx ny krylovdim howmany\n", argv[0]); return 1; }
    const int64_t nx = atoll(argv[1]), ny = atoll(argv[2]);
    const int kd = atoi(argv[3]), howmany = atoi(argv[4]);
    const int64_t n = nx * ny;
    const int tag = B2K_CGS2;
    b2k_ctx* ctx = NULL;
    b2k_op* A = NULL;
    const double c[7] = {4.0, -1.0, -1.0, -1.0, -1.0, 0.0, 0.0};
    b2k_vec* cols = (b2k_vec*)malloc(sizeof(b2k_vec) * (size_t)(kd + 2));
    double* alphas = (double*)calloc((size_t)kd + 1, sizeof(double));
    double* betas = (double*)calloc((size_t)kd + 1, sizeof(double));
    double* D = (double*)malloc(sizeof(double) * (size_t)kd);
    double* U = (double*)malloc(sizeof(double) * (size_t)kd * kd);
    double* f = (double*)malloc(sizeof(double) * (size_t)kd);
    b2k_vec x0, r;
    double beta0, alpha, beta, d;
    int32_t done = 0;
    int K, keep, i, numops;
    if (kd < 4 || kd > 64) { fprintf(stderr, "krylovdim must be in [4, 64]\n"); return 1; }
    CHECK(b2k_ctx_create(&ctx, 0, n, kd + 8, B2K_F64));
    CHECK(b2k_op_create_stencil(ctx, &A, nx, ny, 1, c));
    CHECK(b2k_vec_alloc(ctx, 0, &x0));
    CHECK(b2k_vec_fill_splitmix(ctx, x0, 20260923ull));
    /* initialize */
    CHECK(b2k_vec_norm(ctx, x0, &beta0));
    CHECK(b2k_vec_alloc(ctx, 0, &cols[0]));
    CHECK(b2k_vec_scale(ctx, cols[0], x0, 1.0 / beta0));
    CHECK(b2k_vec_alloc(ctx, 0, &r));
    CHECK(b2k_op_apply_dot(ctx, A, cols[0], r, cols[0], &alpha));
    CHECK(b2k_vec_axpby(ctx, r, cols[0], -alpha, 1.0));
    CHECK(b2k_vec_inner(ctx, cols[0], r, &d));     /* the extra correction of the *2 orthogonalizers, :200-204 */
    alpha += d;
    CHECK(b2k_vec_axpby(ctx, r, cols[0], -d, 1.0));
    CHECK(b2k_vec_norm(ctx, r, &beta));
    alphas[0] = alpha;
    betas[0] = beta;
    K = 1;
    numops = 1;
    /* expand to krylovdim: one call, the steps are chained on the device */
    cols[K] = r;
    CHECK(b2k_lanczos_expand_many(ctx, A, cols, K, kd - K, betas[K - 1], 0.0, tag, 0.0, alphas + K, betas + K, &done, &r));
    K += done;
    numops += done;
    if (K != kd) { fprintf(stderr, "expected %d steps, got %d\n", kd - 1, (int)done); return 3; }
    /* thick restart: keep = (3 krylovdim + 2 converged) / 5 with converged = 0 */
    tridiag_eig(K, alphas, betas, D, U);
    beta = betas[K - 1];
    for (i = 0; i < K; ++i) f[i] = U[(size_t)i * K + (K - 1)] * beta;
    keep = (3 * kd) / 5;
    CHECK(b2k_host_lanczos_restart(K, keep, D, f, U, K, alphas, betas));
    CHECK(b2k_basis_transform(ctx, cols, K, U, K, keep));
    /* B[keep+1] = r / beta (reuses that column), shrink!: columns keep+1 .. K-1 go back to the slab, the popped
     * column is rescaled into the residual (lanczos.jl:277-289) */
    CHECK(b2k_vec_scale(ctx, cols[keep], r, 1.0 / beta));
    CHECK(b2k_vec_free(ctx, r));
    for (i = keep + 1; i < K; ++i) CHECK(b2k_vec_free(ctx, cols[i]));
    r = cols[keep];
    CHECK(b2k_vec_scale(ctx, r, r, betas[keep - 1]));
    K = keep;
    cols[K] = r;
    CHECK(b2k_lanczos_expand_many(ctx, A, cols, K, kd - K, betas[K - 1], 0.0, tag, 0.0, alphas + K, betas + K, &done, &r));
    K += done;
    numops += done;
    tridiag_eig(K, alphas, betas, D, U);
    printf("numops %d\n", numops);
    for (i = 0; i < howmany; ++i) printf("%.17g %.17g\n", D[i], fabs(U[(size_t)i * K + (K - 1)] * betas[K - 1]));
    printf("launches %lld\n", (long long)b2k_ctx_launch_count(ctx));
    free(cols); free(alphas); free(betas); free(D); free(U); free(f);
    CHECK(b2k_op_destroy(ctx, A));
    CHECK(b2k_ctx_destroy(ctx));
    return 0;
}
