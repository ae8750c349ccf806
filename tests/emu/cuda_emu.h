// cuda_emu.h — just enough of the CUDA execution model to run a kernel's SOURCE on host threads (test
// infrastructure; no GPU, no nvcc): one std::thread per CUDA thread of a block, blocks run one after another,
// __syncthreads() = a barrier over the block, __shfl_xor_sync = a slot exchange between two warp-wide barriers,
// dynamic shared memory = a heap buffer per block (so AddressSanitizer sees an out-of-range shared-memory access),
// static __shared__ arrays = function-local statics (safe because blocks do not overlap in time).  Compiled with
// -fsanitize=address or -fsanitize=thread, a run checks a kernel's indexing and its barrier placement.
#pragma once
#include <pthread.h>

#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace b2k_emu {
struct dim3_ { unsigned x = 0, y = 0, z = 0; };
struct Block {
    int nthreads = 0;
    unsigned char* smem = nullptr;
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> warp_bar;
    std::vector<unsigned long long> slots;          // [warp][32]
};
inline thread_local dim3_ t_threadIdx, t_blockIdx, t_gridDim, t_blockDim;
inline thread_local Block* t_block = nullptr;
inline unsigned char* dyn_smem() { return t_block->smem; }
inline void syncthreads() { pthread_barrier_wait(&t_block->bar); }
template <typename T>
inline T shfl_xor(T v, int off) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const int warp = t_threadIdx.x >> 5, lane = t_threadIdx.x & 31;
    unsigned long long bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    t_block->slots[(size_t)warp * 32 + lane] = bits;
    pthread_barrier_wait(&t_block->warp_bar[warp]);
    bits = t_block->slots[(size_t)warp * 32 + (lane ^ off)];
    pthread_barrier_wait(&t_block->warp_bar[warp]);
    T r;
    std::memcpy(&r, &bits, sizeof(T));
    return r;
}
// kernel<<<grid, nthreads, smem_bytes>>>(...): `body` is a closure that calls the kernel with its arguments
inline void launch(int grid, int nthreads, size_t smem_bytes, const std::function<void()>& body) {
    for (int b = 0; b < grid; ++b) {
        Block blk;
        blk.nthreads = nthreads;
        blk.smem = new unsigned char[smem_bytes ? smem_bytes : 1];      // exact size: ASan guards both ends
        const int nwarps = (nthreads + 31) / 32;
        pthread_barrier_init(&blk.bar, nullptr, nthreads);
        blk.warp_bar.resize(nwarps);
        for (int w = 0; w < nwarps; ++w) pthread_barrier_init(&blk.warp_bar[w], nullptr, std::min(32, nthreads - 32 * w));
        blk.slots.assign((size_t)nwarps * 32, 0ull);
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (int t = 0; t < nthreads; ++t)
            th.emplace_back([&, t, b]() {
                t_block = &blk;
                t_threadIdx.x = t; t_blockIdx.x = b; t_gridDim.x = grid; t_blockDim.x = nthreads;
                body();
            });
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&blk.bar);
        for (auto& w : blk.warp_bar) pthread_barrier_destroy(&w);
        delete[] blk.smem;
    }
}
}  // namespace b2k_emu

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __syncthreads() b2k_emu::syncthreads()
#define __shfl_xor_sync(mask, v, off) b2k_emu::shfl_xor((v), (off))
#define threadIdx b2k_emu::t_threadIdx
#define blockIdx b2k_emu::t_blockIdx
#define gridDim b2k_emu::t_gridDim
#define blockDim b2k_emu::t_blockDim
using std::fma;
