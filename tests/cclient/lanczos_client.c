/* Plain-C client of include/b200krylov.h: what a foreign-language binding (Julia ccall, cgo, …) does,
 * without Python in between.  Builds a 5-point Laplacian on the device, runs `steps` Lanczos expansion
 * steps (lanczos.jl:180-222 start, then b2k_lanczos_expand) and prints alpha_k beta_k per step, then the
 * basis orthogonality defect max |<v_i, v_j> - delta_ij|.  tests/test_gpu_cclient.py checks the output
 * against the oracle.  Usage: lanczos_client nx ny steps orth_tag */
#include <stdio.h>
#include <stdlib.h>
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

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s nx ny steps orth_tag\n", argv[0]); return 1; }
    const int64_t nx = atoll(argv[1]), ny = atoll(argv[2]);
    const int steps = atoi(argv[3]), tag = atoi(argv[4]);
    const int64_t n = nx * ny;
    b2k_ctx* ctx = NULL;
    CHECK(b2k_ctx_create(&ctx, 0, n, steps + 6, B2K_F64));
    b2k_op* A = NULL;
    const double c[7] = {4.0, -1.0, -1.0, -1.0, -1.0, 0.0, 0.0};
    CHECK(b2k_op_create_stencil(ctx, &A, nx, ny, 1, c));

    b2k_vec* cols = (b2k_vec*)malloc(sizeof(b2k_vec) * (size_t)(steps + 2));
    b2k_vec x0, r;
    CHECK(b2k_vec_alloc(ctx, 0, &x0));
    CHECK(b2k_vec_fill_splitmix(ctx, x0, 20260923ull));
    /* initialize: v = x0/|x0|, r = A v - alpha v (one extra correction for the *2 orthogonalizers) */
    double beta0, alpha, beta, d;
    CHECK(b2k_vec_norm(ctx, x0, &beta0));
    CHECK(b2k_vec_alloc(ctx, 0, &cols[0]));
    CHECK(b2k_vec_scale(ctx, cols[0], x0, 1.0 / beta0));
    CHECK(b2k_vec_alloc(ctx, 0, &r));
    CHECK(b2k_op_apply_dot(ctx, A, cols[0], r, cols[0], &alpha));
    CHECK(b2k_vec_axpby(ctx, r, cols[0], -alpha, 1.0));
    if (tag == B2K_CGS2 || tag == B2K_MGS2) {
        CHECK(b2k_vec_inner(ctx, cols[0], r, &d));
        alpha += d;
        CHECK(b2k_vec_axpby(ctx, r, cols[0], -d, 1.0));
    }
    CHECK(b2k_vec_norm(ctx, r, &beta));
    printf("%.17g %.17g\n", alpha, beta);
    for (int k = 1; k < steps; ++k) {
        b2k_vec w;
        CHECK(b2k_vec_alloc(ctx, 0, &w));
        cols[k] = r;   /* the residual's storage becomes basis column k+1 */
        CHECK(b2k_lanczos_expand(ctx, A, cols, k, r, w, beta, tag, 0.70710678118654757, &alpha, &beta));
        r = w;
        printf("%.17g %.17g\n", alpha, beta);
    }
    /* orthogonality of the basis through block_inner */
    double* M = (double*)malloc(sizeof(double) * (size_t)steps * (size_t)steps);
    CHECK(b2k_block_inner(ctx, cols, steps, cols, steps, M));
    double defect = 0.0;
    for (int j = 0; j < steps; ++j)
        for (int i = 0; i < steps; ++i) {
            const double e = fabs(M[(size_t)j * steps + i] - (i == j ? 1.0 : 0.0));
            if (e > defect) defect = e;
        }
    printf("defect %.3e launches %lld\n", defect, (long long)b2k_ctx_launch_count(ctx));
    free(M);
    free(cols);
    CHECK(b2k_op_destroy(ctx, A));
    CHECK(b2k_ctx_destroy(ctx));
    return 0;
}
