// dist.cu — NCCL plumbing for row-sharded contexts (one process per GPU).
// NCCL is resolved at run time with dlopen so that (a) the library has no link-time NCCL
// dependency and loads on machines without it, and (b) inside a PyTorch process we bind to
// the very libnccl.so.2 torch already loaded instead of a second copy.
// Collectives per step: AllReduce of <= k+1 doubles (latency-bound) and a
// nearest-neighbour halo exchange for the SpMV.
#include "common.cuh"
#include <dlfcn.h>
#include <cstdlib>
#include <vector>

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclChar_ = 0, ncclFloat64_ = 8 };
enum { ncclSum_ = 0 };

// Small-vector all-reduce over NVLink peer memory (CUDA IPC): every rank publishes its k
// coefficients + a sequence flag in its own buffer, reads the peers' buffers directly over
// NVLink and sums them in rank order (deterministic, identical on all ranks).  One ~5 us
// kernel instead of a ~20-25 us NCCL call for the three latency-bound reductions of a
// Lanczos step.  Falls back to NCCL if IPC mapping is unavailable.
constexpr int PEER_STRIDE = 1024;                 // doubles per slot
constexpr int PEER_MAXK = PEER_STRIDE - 8;
constexpr int PEER_MAXR = 16;
struct PeerArgs {
    double* ptr[PEER_MAXR];     // ptr[p] = rank p's buffer (2 slots + 2 flags), mapped into this process
    int rank, nranks;
};

struct B2kNccl {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
    bool peer_ok = false;
    double* peer_local = nullptr;
    PeerArgs peer;
    unsigned long long peer_seq = 0;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static int32_t load_nccl(b2k_ctx* ctx, B2kNccl* n) {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        n->lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n->lib) break;
    }
    if (!n->lib) return b2k_fail(ctx, B2K_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                   \
    *(void**)(&n->field) = dlsym(n->lib, name);                                            \
    if (!n->field) return b2k_fail(ctx, B2K_ENCCL, "libnccl: missing symbol %s", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllReduce, "ncclAllReduce");
    SYM(AllGather, "ncclAllGather");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return B2K_OK;
}

#define NCCL_CK(ctx, n, call)                                                              \
    do {                                                                                   \
        ncclResult_t r__ = (call);                                                         \
        if (r__ != ncclSuccess_)                                                           \
            return b2k_fail((ctx), B2K_ENCCL, "%s -> %s", #call, (n)->GetErrorString(r__)); \
    } while (0)

extern "C" int32_t b2k_nccl_unique_id(void* uid128) {
    if (!uid128) return B2K_EINVAL;
    B2kNccl n;
    B2K_TRY(load_nccl(nullptr, &n));
    ncclUniqueId id;
    ncclResult_t r = n.GetUniqueId(&id);
    if (r != ncclSuccess_) return b2k_fail(nullptr, B2K_ENCCL, "ncclGetUniqueId -> %s", n.GetErrorString(r));
    memcpy(uid128, &id, sizeof(id));
    return B2K_OK;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_volatile_f64(const double* p) {
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
k_peer_allreduce(const __grid_constant__ PeerArgs a, double* __restrict__ inout, int count,
                 unsigned long long seq) {
    const int slot = (int)(seq & 1ull);
    double* mine = a.ptr[a.rank] + (size_t)slot * PEER_STRIDE;
    unsigned long long* myflag = reinterpret_cast<unsigned long long*>(a.ptr[a.rank] + 2 * PEER_STRIDE) + slot;
    for (int i = threadIdx.x; i < count; i += blockDim.x) mine[i] = inout[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(myflag, seq);
    if (threadIdx.x < a.nranks && threadIdx.x != a.rank) {
        const unsigned long long* pf =
            reinterpret_cast<const unsigned long long*>(a.ptr[threadIdx.x] + 2 * PEER_STRIDE) + slot;
        while (ld_acquire_sys(pf) < seq) {
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double acc = 0.0;
        for (int p = 0; p < a.nranks; ++p)     // rank order: every rank computes the same bits
            acc += ld_volatile_f64(a.ptr[p] + (size_t)slot * PEER_STRIDE + i);
        inout[i] = acc;
    }
}

static void peer_setup(b2k_ctx* ctx, B2kNccl* n) {
    // NOTE: every rank runs the same collectives below, whatever its local outcome.
    n->peer_ok = false;
    // opt-in (B2K_PEER=1): validated on 2 GPUs in r01 (bit-identical to the NCCL path, no speed-up
    // at N=2); NCCL stays the default until it is measured at 8 GPUs.
    const char* on = getenv("B2K_PEER");
    if (!(on && on[0] == '1') || ctx->nranks > PEER_MAXR) return;   // same decision on every rank
    const int R = ctx->nranks;
    const size_t bytes = sizeof(double) * (2 * PEER_STRIDE + 16);
    int good = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&n->peer_local, bytes) != cudaSuccess) { cudaGetLastError(); n->peer_local = nullptr; good = 0; }
    if (good) cudaMemset(n->peer_local, 0, bytes);
    if (good && cudaIpcGetMemHandle(&mine, n->peer_local) != cudaSuccess) { cudaGetLastError(); good = 0; }
    char* d_h = nullptr;
    std::vector<cudaIpcMemHandle_t> all(R);
    if (cudaMalloc(&d_h, sizeof(mine) * R + sizeof(double)) != cudaSuccess) return;   // cannot even talk
    cudaMemcpy(d_h + sizeof(mine) * ctx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    if (n->AllGather(d_h + sizeof(mine) * ctx->rank, d_h, sizeof(mine), ncclChar_, n->comm, ctx->stream) != ncclSuccess_) good = 0;
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(all.data(), d_h, sizeof(mine) * R, cudaMemcpyDeviceToHost);
    for (int p = 0; p < R && good; ++p) {
        if (p == ctx->rank) { n->peer.ptr[p] = n->peer_local; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[p], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            break;
        }
        n->peer.ptr[p] = (double*)q;
    }
    // every rank must agree (and have finished mapping) before the first use
    double* d_flag = reinterpret_cast<double*>(d_h + sizeof(mine) * R);
    double hv = good ? 0.0 : 1.0;
    cudaMemcpy(d_flag, &hv, sizeof(double), cudaMemcpyHostToDevice);
    n->AllReduce(d_flag, d_flag, 1, ncclFloat64_, ncclSum_, n->comm, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(&hv, d_flag, sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d_h);
    n->peer.rank = ctx->rank;
    n->peer.nranks = R;
    n->peer_ok = (hv == 0.0);
    n->peer_seq = 0;
}

bool b2k_peer_ok(const b2k_ctx* ctx) { return ctx->nccl && ctx->nccl->peer_ok; }

int32_t b2k_peer_allreduce(b2k_ctx* ctx, double* dptr, int32_t count) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->peer_ok || count > PEER_MAXK) return b2k_fail(ctx, B2K_ENCCL, "peer all-reduce unavailable");
    ++n->peer_seq;
    k_peer_allreduce<<<1, 256, 0, ctx->stream>>>(n->peer, dptr, count, n->peer_seq);
    B2K_LAUNCH_CHECK(ctx);
    return B2K_OK;
}

// One NCCL communicator per process (rank, nranks, device) is kept alive and shared by
// consecutive contexts: ncclCommInitRank costs 50-200 ms, which would dominate a host-buffer
// solve that creates a context per call.  A cached communicator ignores the new unique id.
static B2kNccl* g_comm_cache = nullptr;
static int g_comm_rank = -1, g_comm_nranks = -1, g_comm_device = -1;
static bool g_comm_in_use = false;

int32_t b2k_nccl_init(b2k_ctx* ctx, const void* uid) {
    if (g_comm_cache && !g_comm_in_use && g_comm_rank == ctx->rank && g_comm_nranks == ctx->nranks &&
        g_comm_device == ctx->device) {
        ctx->nccl = g_comm_cache;
        g_comm_in_use = true;
        return B2K_OK;
    }
    if (!uid) return b2k_fail(ctx, B2K_EINVAL, "ctx_create_dist: nccl_uid is NULL");
    B2kNccl* n = new B2kNccl();
    int32_t rc = load_nccl(ctx, n);
    if (rc != B2K_OK) {
        delete n;
        return rc;
    }
    ncclUniqueId id;
    memcpy(&id, uid, sizeof(id));
    ctx->nccl = n;
    NCCL_CK(ctx, n, n->CommInitRank(&n->comm, ctx->nranks, id, ctx->rank));
    peer_setup(ctx, n);
    if (!g_comm_cache) {
        g_comm_cache = n;
        g_comm_rank = ctx->rank;
        g_comm_nranks = ctx->nranks;
        g_comm_device = ctx->device;
        g_comm_in_use = true;
    }
    return B2K_OK;
}

void b2k_nccl_destroy(b2k_ctx* ctx) {
    if (!ctx->nccl) return;
    if (ctx->nccl == g_comm_cache) {     // stays alive for the next context of this process
        g_comm_in_use = false;
        ctx->nccl = nullptr;
        return;
    }
    if (ctx->nccl->peer_ok)
        for (int p = 0; p < ctx->nranks; ++p)
            if (p != ctx->rank) cudaIpcCloseMemHandle(ctx->nccl->peer.ptr[p]);
    if (ctx->nccl->peer_local) cudaFree(ctx->nccl->peer_local);
    if (ctx->nccl->comm) ctx->nccl->CommDestroy(ctx->nccl->comm);
    delete ctx->nccl;
    ctx->nccl = nullptr;
}

int32_t b2k_nccl_allreduce_f64(b2k_ctx* ctx, double* dptr, int32_t count) {
    B2kNccl* n = ctx->nccl;
    if (!n) return b2k_fail(ctx, B2K_ENCCL, "allreduce on a context without a communicator");
    NCCL_CK(ctx, n, n->AllReduce(dptr, dptr, (size_t)count, ncclFloat64_, ncclSum_, n->comm, ctx->stream));
    return B2K_OK;
}

int32_t b2k_nccl_allgather(b2k_ctx* ctx, const void* sendbuf, void* recvbuf, size_t bytes) {
    B2kNccl* n = ctx->nccl;
    if (!n) return b2k_fail(ctx, B2K_ENCCL, "allgather on a context without a communicator");
    NCCL_CK(ctx, n, n->AllGather(sendbuf, recvbuf, bytes, ncclChar_, n->comm, ctx->stream));
    return B2K_OK;
}

int32_t b2k_nccl_halo_exchange(b2k_ctx* ctx, int up, int dn, const void* send_up, size_t send_up_bytes,
                               void* recv_dn, size_t recv_dn_bytes, const void* send_dn,
                               size_t send_dn_bytes, void* recv_up, size_t recv_up_bytes) {
    B2kNccl* n = ctx->nccl;
    if (!n) return b2k_fail(ctx, B2K_ENCCL, "halo exchange on a context without a communicator");
    NCCL_CK(ctx, n, n->GroupStart());
    if (up >= 0 && send_up_bytes) NCCL_CK(ctx, n, n->Send(send_up, send_up_bytes, ncclChar_, up, n->comm, ctx->stream));
    if (dn >= 0 && recv_dn_bytes) NCCL_CK(ctx, n, n->Recv(recv_dn, recv_dn_bytes, ncclChar_, dn, n->comm, ctx->stream));
    if (dn >= 0 && send_dn_bytes) NCCL_CK(ctx, n, n->Send(send_dn, send_dn_bytes, ncclChar_, dn, n->comm, ctx->stream));
    if (up >= 0 && recv_up_bytes) NCCL_CK(ctx, n, n->Recv(recv_up, recv_up_bytes, ncclChar_, up, n->comm, ctx->stream));
    NCCL_CK(ctx, n, n->GroupEnd());
    return B2K_OK;
}
