// Probe: what limits a plain streaming kernel (y = a*x, 80 MB + 80 MB) on B200?  Variants of grid shape, occupancy,
// unrolling and cache hints, timed with CUDA events; prints GB/s for each.  nvcc -arch=sm_100a -O3 -o blas1_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int U, int HINT>
__global__ void k_persist(double* __restrict__ y, const double* __restrict__ x, long nv, double a) {
    const long stride = (long)gridDim.x * blockDim.x;
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < nv; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const double2* p = reinterpret_cast<const double2*>(x) + i + u * stride;
            v[u] = HINT ? __ldcs(p) : __ldg(p);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u].x *= a; v[u].y *= a;
            double2* q = reinterpret_cast<double2*>(y) + i + u * stride;
            if (HINT) __stcs(q, v[u]); else *q = v[u];
        }
    }
    for (; i < nv; i += stride) {
        double2 v = reinterpret_cast<const double2*>(x)[i];
        v.x *= a; v.y *= a;
        reinterpret_cast<double2*>(y)[i] = v;
    }
}

// non-persistent: each CTA owns a contiguous chunk of U*blockDim vectors
template <int U, int HINT>
__global__ void k_chunk(double* __restrict__ y, const double* __restrict__ x, long nv, double a) {
    const long base = (long)blockIdx.x * (U * blockDim.x) + threadIdx.x;
    double2 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long i = base + (long)u * blockDim.x;
        if (i < nv) { const double2* p = reinterpret_cast<const double2*>(x) + i; v[u] = HINT ? __ldcs(p) : __ldg(p); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long i = base + (long)u * blockDim.x;
        if (i < nv) { v[u].x *= a; v[u].y *= a; double2* q = reinterpret_cast<double2*>(y) + i; if (HINT) __stcs(q, v[u]); else *q = v[u]; }
    }
}

template <typename F>
float timeit(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const long n = 10000000, nv = n / 2;
    double *x, *y, *big;
    CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8)); CK(cudaMalloc(&big, 512l << 20));
    CK(cudaMemset(x, 0, n * 8)); CK(cudaMemset(y, 0, n * 8));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double bytes = 2.0 * n * 8;
    auto rep = [&](const char* name, float ms) { printf("%-44s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
    // L2 is 126 MB and x, y are 80 MB each: alternate with a 512 MB memset so that every rep starts cold? No: the
    // product kernels run back to back on 5+ GB working sets; emulate that by striding over `big` between reps
    auto flush = [&]() { cudaMemsetAsync(big, 1, 512l << 20); };
    float base = timeit([&]() { flush(); }, 10);
    printf("flush alone %.2f us\n", base * 1e3);
#define RUN(name, launch) rep(name, timeit([&]() { flush(); launch; }, 10) - base)
    RUN("persist 4/SM x256 U1", (k_persist<1, 0><<<sms * 4, 256>>>(y, x, nv, 0.5)));
    RUN("persist 4/SM x256 U4", (k_persist<4, 0><<<sms * 4, 256>>>(y, x, nv, 0.5)));
    RUN("persist 8/SM x256 U4", (k_persist<4, 0><<<sms * 8, 256>>>(y, x, nv, 0.5)));
    RUN("persist 8/SM x256 U4 cs", (k_persist<4, 1><<<sms * 8, 256>>>(y, x, nv, 0.5)));
    RUN("persist 2/SM x1024 U4", (k_persist<4, 0><<<sms * 2, 1024>>>(y, x, nv, 0.5)));
    RUN("persist 16/SM x128 U8", (k_persist<8, 0><<<sms * 16, 128>>>(y, x, nv, 0.5)));
    RUN("chunk x256 U4", (k_chunk<4, 0><<<(unsigned)((nv + 1023) / 1024), 256>>>(y, x, nv, 0.5)));
    RUN("chunk x256 U4 cs", (k_chunk<4, 1><<<(unsigned)((nv + 1023) / 1024), 256>>>(y, x, nv, 0.5)));
    RUN("chunk x256 U8", (k_chunk<8, 0><<<(unsigned)((nv + 2047) / 2048), 256>>>(y, x, nv, 0.5)));
    RUN("chunk x512 U4", (k_chunk<4, 0><<<(unsigned)((nv + 2047) / 2048), 512>>>(y, x, nv, 0.5)));
    RUN("chunk x128 U8 cs", (k_chunk<8, 1><<<(unsigned)((nv + 1023) / 1024), 128>>>(y, x, nv, 0.5)));
    RUN("cudaMemcpyAsync D2D", cudaMemcpyAsync(y, x, n * 8, cudaMemcpyDeviceToDevice));
    // no flush: L2-warm back-to-back (what tools/microbench.py measures)
    rep("persist 4/SM x256 U4 (no flush, warm)", timeit([&]() { k_persist<4, 0><<<sms * 4, 256>>>(y, x, nv, 0.5); }, 20));
    rep("chunk x256 U4 (no flush, warm)", timeit([&]() { k_chunk<4, 0><<<(unsigned)((nv + 1023) / 1024), 256>>>(y, x, nv, 0.5); }, 20));
    return 0;
}
