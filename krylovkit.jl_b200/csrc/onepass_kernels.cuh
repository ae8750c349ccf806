// onepass_kernels.cuh — the kernels of the one-pass dense Golub-Kahan-Lanczos step (see spmv.cu for the design
// notes and the host side).  The file is included by spmv.cu inside its anonymous namespace, and — with
// B2K_HOST_EMU defined — by tests/emu/onepass_emu.cpp, which compiles THESE SAME kernel bodies with g++ over a small
// CUDA-on-threads emulation (one std::thread per CUDA thread, barriers for __syncthreads, slot exchange for the warp
// shuffles) so that their index arithmetic, synchronisation and reductions run on a box without a GPU.  Only the
// 16-byte streaming loads (inline PTX) and the dynamic shared-memory declaration differ between the two builds.
#pragma once

#ifdef B2K_HOST_EMU
#define B2K_DYN_SMEM(name) unsigned char* name = b2k_emu::dyn_smem()
#else
#define B2K_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

constexpr int OP_T = 256;        // threads per CTA
constexpr int OP_ROWS = 32;      // rows per tile
constexpr int OP_PAD = 33;       // column stride of the tile in shared memory (words)
constexpr int OP_ZMAX = 7;       // columns per thread in phase 2: n <= OP_ZMAX * OP_T
constexpr int OP_UNR = 8;        // 16-byte loads in flight per thread

template <typename T> struct OpVecT;
template <> struct OpVecT<float>  { using type = float4; };
template <> struct OpVecT<double> { using type = double2; };
__device__ __forceinline__ void op_unpack(const float4& a, float* e)  { e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; }
__device__ __forceinline__ void op_unpack(const double2& a, double* e) { e[0] = a.x; e[1] = a.y; }
#ifdef B2K_HOST_EMU
inline float4 op_ld_stream(const float4* p) { return *p; }
inline double2 op_ld_stream(const double2* p) { return *p; }
#else
__device__ __forceinline__ float4 op_ld_stream(const float4* p) {        // read-once stream: no L1 allocation
    float4 a;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    return a;
}
__device__ __forceinline__ double2 op_ld_stream(const double2* p) {
    double2 a;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(a.x), "=d"(a.y) : "l"(p));
    return a;
}
#endif

template <typename T, int NZ>               // NZ = columns per thread in phase 2: n <= NZ * OP_T
__global__ void __launch_bounds__(OP_T, NZ <= 2 ? 3 : 1)
k_dense_onepass(const T* __restrict__ A, int64_t ld, int64_t m, int32_t n, const T* __restrict__ x,
                T* __restrict__ y, double* __restrict__ zpart, int64_t ntiles) {
    using V = typename OpVecT<T>::type;
    constexpr int VEC = 16 / (int)sizeof(T);        // rows per 16-byte load
    constexpr int VPC = OP_ROWS / VEC;              // loads per tile column
    constexpr int CSTEP = OP_T / VPC;               // columns covered by one load of the whole CTA
    static_assert(OP_T % VPC == 0 && 32 % VPC == 0, "a thread keeps the same rows for every column it loads");
    B2K_DYN_SMEM(op_smem);
    T* As = reinterpret_cast<T*>(op_smem);          // [n][OP_PAD]
    T* xs = As + (size_t)n * OP_PAD;                // [n]
    T* ys_part = xs + n;                            // [8 warps][32 rows]
    T* ys = ys_part + (OP_T / 32) * OP_ROWS;        // [32 rows]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rq = tid % VPC;                       // which 16-byte row group of a column this thread loads
    const int c0 = tid / VPC;                       // its first column; the others follow at CSTEP
    for (int j = tid; j < n; j += OP_T) xs[j] = x[j];
    double zacc[NZ];
#pragma unroll
    for (int s = 0; s < NZ; ++s) zacc[s] = 0.0;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * OP_ROWS;
        const T* At = A + i0 + rq * VEC;
        T acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = (T)0;
        for (int cb = c0; cb < n; cb += CSTEP * OP_UNR) {
            V a[OP_UNR];
#pragma unroll
            for (int u = 0; u < OP_UNR; ++u) {
                const int c = cb + u * CSTEP;
                if (c < n) a[u] = op_ld_stream(reinterpret_cast<const V*>(At + (int64_t)c * ld));
            }
#pragma unroll
            for (int u = 0; u < OP_UNR; ++u) {
                const int c = cb + u * CSTEP;
                if (c < n) {
                    T e4[VEC];
                    op_unpack(a[u], e4);
                    const T xc = xs[c];
                    T* dst = As + (size_t)c * OP_PAD + rq * VEC;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        acc[e] = fma(e4[e], xc, acc[e]);
                        dst[e] = e4[e];
                    }
                }
            }
        }
        // the threads of a warp that hold the same rows: lanes rq, rq + VPC, rq + 2 VPC, ...
#pragma unroll
        for (int off = VPC; off < 32; off <<= 1)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        if (lane < VPC)
#pragma unroll
            for (int e = 0; e < VEC; ++e) ys_part[warp * OP_ROWS + lane * VEC + e] = acc[e];
        __syncthreads();                            // the tile and the per-warp partial y are complete
        if (tid < OP_ROWS) {
            T s = (T)0;
#pragma unroll
            for (int w = 0; w < OP_T / 32; ++w) s += ys_part[w * OP_ROWS + tid];
            ys[tid] = s;
            if (i0 + tid < m) y[i0 + tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int s = 0; s < NZ; ++s) {
            const int c = s * OP_T + tid;
            if (c < n) {
                const T* col = As + (size_t)c * OP_PAD;
                T p = (T)0;
#pragma unroll
                for (int row = 0; row < OP_ROWS; ++row) p = fma(col[row], ys[row], p);
                zacc[s] += (double)p;
            }
        }
        __syncthreads();                            // phase 2 has read the tile: the next one may overwrite it
    }
#pragma unroll
    for (int s = 0; s < NZ; ++s) {
        const int c = s * OP_T + tid;
        if (c < n) zpart[(size_t)blockIdx.x * n + c] = zacc[s];
    }
}

// Variant B of the same step for the shape config 4 has (Float32, n <= 512): 64-row tiles — 256-byte column segments
// instead of 128 — in a 512-thread CTA that owns the SM (139 KB of shared memory), with the overlap a single CTA
// per SM cannot get from its neighbours made explicit: every thread keeps its 16 loads of the NEXT tile in flight
// in registers while the CTA reduces y and forms z for the current one; a register is refilled from the next tile
// as soon as its value has been parked in shared memory.  Column stride 65 words (lane <-> column reads in phase 2
// are conflict-free; the stores of 16 row groups x 2 columns per warp are two-way conflicted).  Selected by
// b2k_debug_set_onepass_variant(1); the default stays variant A until both have been measured on a B200.
constexpr int OPW_T = 512;
constexpr int OPW_ROWS = 64;
constexpr int OPW_PAD = 65;
constexpr int OPW_LD = 16;       // loads per thread per tile: n <= OPW_LD * (OPW_T / 16) = 512

__global__ void __launch_bounds__(OPW_T, 1)
k_dense_onepass_w(const float* __restrict__ A, int64_t ld, int64_t m, int32_t n, const float* __restrict__ x,
                  float* __restrict__ y, double* __restrict__ zpart, int64_t ntiles) {
    constexpr int VPC = OPW_ROWS / 4;               // 16 float4 per tile column
    constexpr int CSTEP = OPW_T / VPC;              // 32 columns per load of the whole CTA
    B2K_DYN_SMEM(op_smem);
    float* As = reinterpret_cast<float*>(op_smem);  // [n][OPW_PAD]
    float* xs = As + (size_t)n * OPW_PAD;           // [n]
    float* ys_part = xs + n;                        // [16 warps][64 rows]
    float* ys = ys_part + (OPW_T / 32) * OPW_ROWS;  // [64 rows]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rq = tid % VPC;
    const int c0 = tid / VPC;
    for (int j = tid; j < n; j += OPW_T) xs[j] = x[j];
    double zacc = 0.0;
    float4 a[OPW_LD];
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // rows [ld, ...) do not exist (ld is a multiple of 32, a tile has 64 rows): such a row group loads as zero
    {
        const int64_t r0 = (int64_t)blockIdx.x * OPW_ROWS + rq * 4;
        const bool rows_ok = (int64_t)blockIdx.x < ntiles && r0 < ld;
        const float* At = A + r0;
#pragma unroll
        for (int u = 0; u < OPW_LD; ++u) {
            const int c = c0 + u * CSTEP;
            a[u] = (rows_ok && c < n) ? op_ld_stream(reinterpret_cast<const float4*>(At + (int64_t)c * ld)) : zero4;
        }
    }
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * OPW_ROWS;
        const int64_t next = tile + gridDim.x;
        const int64_t r0n = next * OPW_ROWS + rq * 4;
        const bool next_ok = next < ntiles && r0n < ld;
        const float* Atn = A + r0n;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < OPW_LD; ++u) {
            const int c = c0 + u * CSTEP;
            if (c < n) {
                float e4[4];
                op_unpack(a[u], e4);
                const float xc = xs[c];
                float* dst = As + (size_t)c * OPW_PAD + rq * 4;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc[e] = fmaf(e4[e], xc, acc[e]);
                    dst[e] = e4[e];
                }
                // the register is free again: refill it from the next tile of this CTA
                a[u] = next_ok ? op_ld_stream(reinterpret_cast<const float4*>(Atn + (int64_t)c * ld)) : zero4;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        if (lane < VPC)
#pragma unroll
            for (int e = 0; e < 4; ++e) ys_part[warp * OPW_ROWS + lane * 4 + e] = acc[e];
        __syncthreads();
        if (tid < OPW_ROWS) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < OPW_T / 32; ++w) s += ys_part[w * OPW_ROWS + tid];
            ys[tid] = s;
            if (i0 + tid < m) y[i0 + tid] = s;
        }
        __syncthreads();
        if (tid < n) {
            const float* col = As + (size_t)tid * OPW_PAD;
            float p0 = 0.f, p1 = 0.f;
#pragma unroll
            for (int row = 0; row < OPW_ROWS; row += 2) {
                p0 = fmaf(col[row], ys[row], p0);
                p1 = fmaf(col[row + 1], ys[row + 1], p1);
            }
            zacc += (double)(p0 + p1);
        }
        __syncthreads();
    }
    if (tid < n) zpart[(size_t)blockIdx.x * n + tid] = zacc;
}

// z[c] = sum over the CTAs' partials, in a fixed order: 8 groups of threads take every 8th partial with four
// independent running sums each (loads batched), then the groups are added in order.
template <typename T>
__global__ void __launch_bounds__(256)
k_onepass_reduce(const double* __restrict__ zpart, int nparts, int n, double* __restrict__ dres, T* __restrict__ zout) {
    __shared__ double red[8][32];
    const int g = threadIdx.x >> 5, jl = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + jl;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (c < n) {
        int p = g;
        for (; p + 24 < nparts; p += 32) {
            const double a0 = zpart[(size_t)p * n + c], a1 = zpart[(size_t)(p + 8) * n + c];
            const double a2 = zpart[(size_t)(p + 16) * n + c], a3 = zpart[(size_t)(p + 24) * n + c];
            s0 += a0; s1 += a1; s2 += a2; s3 += a3;
        }
        for (; p < nparts; p += 8) s0 += zpart[(size_t)p * n + c];
    }
    red[g][jl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (g == 0 && c < n) {
        double s = red[0][jl];
#pragma unroll
        for (int w = 1; w < 8; ++w) s += red[w][jl];
        dres[c] = s;
        if (zout) zout[c] = (T)s;
    }
}

template <typename T>
__global__ void k_onepass_store(const double* __restrict__ dres, T* __restrict__ zout, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n) zout[c] = (T)dres[c];
}

