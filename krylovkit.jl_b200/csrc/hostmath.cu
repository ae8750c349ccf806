// hostmath.cu — small dense k x k host algebra that KrylovKit does in compiled Julia code
// (src/dense/reflector.jl, the restart block of src/eigsolve/lanczos.jl:88-105).  Pure host
// C++ (no device code): the Python mirror would otherwise spend ~3 ms of interpreter time
// per restart in this O(keep^2 K) loop while the GPU idles.
#include <cmath>
#include <cstdint>
#include <vector>
#include "../../include/b200krylov.h"

namespace {

// _householder!(v, i) — dense/reflector.jl:34-65 (real): returns beta, nu; v overwritten
void householder_vec(double* v, int len, int i, double* beta, double* nu) {
    double sigma = 0.0;
    for (int k = 0; k < i; ++k) sigma += v[k] * v[k];
    for (int k = i + 1; k < len; ++k) sigma += v[k] * v[k];
    double vi = v[i];
    const double n = std::sqrt(vi * vi + sigma);
    *nu = n;
    if (sigma == 0.0 && vi == n) {
        *beta = 0.0;
        return;
    }
    if (vi < 0) vi = vi - n;
    else vi = -sigma / (vi + n);
    for (int k = 0; k < i; ++k) v[k] /= vi;
    for (int k = i + 1; k < len; ++k) v[k] /= vi;
    v[i] = 1.0;
    *beta = -vi / n;
}

}  // namespace

// Restore Lanczos (tridiagonal) form in the first `keep` columns after a thick restart —
// src/eigsolve/lanczos.jl:88-105.  D: sorted Ritz values (>= keep entries), f: residual weights,
// U: K x K column-major eigenvector matrix (ldu), updated in place by rmul!(U, h').
// alphas/betas receive H[j,j] and H[j+1,j], j < keep.
extern "C" int32_t b2k_host_lanczos_restart(int32_t K, int32_t keep, const double* D, const double* f,
                                            double* U, int32_t ldu, double* alphas, double* betas) {
    if (K < 1 || keep < 1 || keep >= K + 1 || !D || !f || !U || ldu < K || !alphas || !betas) return B2K_EINVAL;
    const int ldh = keep + 1;
    std::vector<double> H((size_t)ldh * keep, 0.0), v(keep), w(std::max(K, ldh));
#define HH(i, j) H[(size_t)(j) * ldh + (i)]
    for (int j = 0; j < keep; ++j) {
        HH(j, j) = D[j];
        HH(keep, j) = f[j];
    }
    for (int j = keep - 1; j >= 0; --j) {
        // h, nu = householder(H, j+1, 1:j, j)   (row form, reflector.jl:24-29)
        const int len = j + 1;
        for (int c = 0; c < len; ++c) v[c] = HH(j + 1, c);
        double beta, nu;
        householder_vec(v.data(), len, j, &beta, &nu);
        HH(j + 1, j) = nu;
        for (int c = 0; c < j; ++c) HH(j + 1, c) = 0.0;
        if (beta == 0.0) continue;
        // lmul!(h, H): rows 0..j of every column
        for (int c = 0; c < keep; ++c) {
            double mu = 0.0;
            for (int r = 0; r < len; ++r) mu += v[r] * HH(r, c);
            mu *= beta;
            for (int r = 0; r < len; ++r) HH(r, c) -= mu * v[r];
        }
        // rmul!(view(H, 1:j, :), h'): columns 0..j, rows 0..j
        for (int r = 0; r < len; ++r) w[r] = 0.0;
        for (int c = 0; c < len; ++c)
            for (int r = 0; r < len; ++r) w[r] += HH(r, c) * v[c];
        for (int c = 0; c < len; ++c) {
            const double vb = beta * v[c];
            for (int r = 0; r < len; ++r) HH(r, c) -= w[r] * vb;
        }
        // rmul!(U, h'): columns 0..j of U, all K rows
        for (int r = 0; r < K; ++r) w[r] = 0.0;
        for (int c = 0; c < len; ++c)
            for (int r = 0; r < K; ++r) w[r] += U[(size_t)c * ldu + r] * v[c];
        for (int c = 0; c < len; ++c) {
            const double vb = beta * v[c];
            for (int r = 0; r < K; ++r) U[(size_t)c * ldu + r] -= w[r] * vb;
        }
    }
    for (int j = 0; j < keep; ++j) {
        alphas[j] = HH(j, j);
        betas[j] = HH(j + 1, j);
    }
#undef HH
    return B2K_OK;
}
