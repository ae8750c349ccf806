// Probe: multi-stream elementwise kernels on arrays far larger than L2 (NV doubles each) — the shapes of the CG /
// BiCGStab updates (2 reads + 1 write, 4 reads + 2 writes) — as a persistent grid-stride kernel (the library's launch
// shape, U vectors per trip) and as a non-persistent one-chunk-per-CTA kernel.  Prints GB/s of algorithmic bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// NR read-only streams, NW read-modify-write streams
template <int NR, int NW, int U>
__global__ void __launch_bounds__(256) k_persist(double* const* __restrict__ w, const double* const* __restrict__ r, long nv, double a) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x; i0 < nv; i0 += U * stride) {
        double2 rv[U][NR], wv[U][NW];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = i0 + u * stride;
            if (i < nv) {
#pragma unroll
                for (int k = 0; k < NR; ++k) rv[u][k] = reinterpret_cast<const double2*>(r[k])[i];
#pragma unroll
                for (int k = 0; k < NW; ++k) wv[u][k] = reinterpret_cast<const double2*>(w[k])[i];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long i = i0 + u * stride;
            if (i < nv) {
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    double2 v = wv[u][k];
#pragma unroll
                    for (int j = 0; j < NR; ++j) { v.x = fma(a, rv[u][j].x, v.x); v.y = fma(a, rv[u][j].y, v.y); }
                    reinterpret_cast<double2*>(w[k])[i] = v;
                }
            }
        }
    }
}

template <int NR, int NW, int U>
__global__ void __launch_bounds__(256) k_chunk(double* const* __restrict__ w, const double* const* __restrict__ r, long nv, double a) {
    const long base = (long)blockIdx.x * (U * blockDim.x) + threadIdx.x;
    double2 rv[U][NR], wv[U][NW];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long i = base + (long)u * blockDim.x;
        if (i < nv) {
#pragma unroll
            for (int k = 0; k < NR; ++k) rv[u][k] = reinterpret_cast<const double2*>(r[k])[i];
#pragma unroll
            for (int k = 0; k < NW; ++k) wv[u][k] = reinterpret_cast<const double2*>(w[k])[i];
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const long i = base + (long)u * blockDim.x;
        if (i < nv) {
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                double2 v = wv[u][k];
#pragma unroll
                for (int j = 0; j < NR; ++j) { v.x = fma(a, rv[u][j].x, v.x); v.y = fma(a, rv[u][j].y, v.y); }
                reinterpret_cast<double2*>(w[k])[i] = v;
            }
        }
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

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 100000000l;     // doubles per array (800 MB)
    const long nv = n / 2;
    double* buf[6];
    const long pad = argc > 2 ? atol(argv[2]) : -1;           // >= 0: one allocation, columns (n + pad) doubles apart
    if (pad >= 0) {
        double* slab;
        CK(cudaMalloc(&slab, 6 * (n + pad) * 8));
        CK(cudaMemset(slab, 0, 6 * (n + pad) * 8));
        for (int i = 0; i < 6; ++i) buf[i] = slab + i * (n + pad);
        printf("slab layout, column pitch %ld doubles\n", n + pad);
    } else {
        for (int i = 0; i < 6; ++i) { CK(cudaMalloc(&buf[i], n * 8)); CK(cudaMemset(buf[i], 0, n * 8)); }
    }
    printf("n = %ld doubles per array (%.0f MB)\n", n, n * 8 / 1e6);
    double** dptr;
    CK(cudaMalloc(&dptr, 6 * sizeof(double*)));
    CK(cudaMemcpy(dptr, buf, 6 * sizeof(double*), cudaMemcpyHostToDevice));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    if (argc > 3) {                                           // persisting-L2 set-aside, like the library's contexts
        cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
        const size_t want = (size_t)atol(argv[3]) < (size_t)prop.persistingL2CacheMaxSize ? (size_t)atol(argv[3]) : (size_t)prop.persistingL2CacheMaxSize;
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        printf("persisting L2 set-aside %zu bytes (device max %d, L2 %d)\n", want, prop.persistingL2CacheMaxSize, prop.l2CacheSize);
    }
    // warm the clocks
    timeit([&]() { k_persist<1, 1, 4><<<sms * 4, 256>>>(dptr, (const double* const*)(dptr + 2), nv, 0.0); }, 30);
#define RUN(NR, NW, name, launch) { float ms = timeit([&]() { launch; }, 10); printf("%-46s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, (NR + 2.0 * NW) * n * 8 / ms / 1e6); }
#define PERSIST(NR, NW, U, PER) RUN(NR, NW, "persist " #NR "r+" #NW "rw U" #U " " #PER "/SM", (k_persist<NR, NW, U><<<sms * PER, 256>>>(dptr, (const double* const*)(dptr + 2), nv, 0.0)))
#define CHUNK(NR, NW, U) RUN(NR, NW, "chunk   " #NR "r+" #NW "rw U" #U, (k_chunk<NR, NW, U><<<(unsigned)((nv + U * 256 - 1) / (U * 256)), 256>>>(dptr, (const double* const*)(dptr + 2), nv, 0.0)))
    PERSIST(1, 1, 1, 4) PERSIST(1, 1, 4, 4) PERSIST(1, 1, 4, 8) PERSIST(1, 1, 8, 4) CHUNK(1, 1, 4) CHUNK(1, 1, 8)
    PERSIST(2, 2, 1, 4) PERSIST(2, 2, 2, 4) PERSIST(2, 2, 4, 4) PERSIST(2, 2, 2, 8) CHUNK(2, 2, 2) CHUNK(2, 2, 4)
    PERSIST(3, 2, 2, 4) CHUNK(3, 2, 2) CHUNK(3, 2, 4)
    RUN(1, 0.5, "cudaMemcpyAsync D2D (800 MB)", cudaMemcpyAsync(buf[0], buf[2], n * 8, cudaMemcpyDeviceToDevice));
    return 0;
}
