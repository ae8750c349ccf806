// dist.cu — transports of row-sharded contexts (one process per GPU): the NVLink peer window (CUDA IPC, the
// default for every latency-bound exchange of the Krylov step) and NCCL (bootstrap-free fallback, bulk gathers).
// NCCL is resolved at run time with dlopen so that (a) the library has no link-time NCCL
// dependency and loads on machines without it, and (b) inside a PyTorch process we bind to
// the very libnccl.so.2 torch already loaded instead of a second copy.
// Collectives per step: AllReduce of <= k+1 doubles (latency-bound) and a
// nearest-neighbour halo exchange for the SpMV.
#include "common.cuh"
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cstdlib>
#include <ctime>
#include <vector>

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclChar_ = 0, ncclFloat64_ = 8 };
enum { ncclSum_ = 0 };

// Node-local rendezvous: a POSIX shared-memory segment named after the job's 128-byte unique id.  It carries
// the CUDA IPC handles (and other small blobs, b2k_host_allgather) between the ranks of one node, so the peer
// window can be set up without NCCL — and with it, two ranks can share ONE GPU (NCCL refuses that), which is
// how the sharded path is tested on a single-GPU box.
constexpr size_t RV_BLOB = 4096;
struct RvShm {
    unsigned long long arrive;
    unsigned long long pad[7];
    unsigned char data[2][PEER_MAXR][RV_BLOB];
};
struct Rendezvous {
    RvShm* shm = nullptr;
    char name[64] = {0};
    unsigned long long round = 0;
    int rank = 0, nranks = 1;
};

struct B2kNccl {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;          // nullptr: NCCL not in use (B2K_NO_NCCL=1 or libnccl missing)
    Rendezvous rv;
    bool peer_ok = false;
    char* win_local = nullptr;          // my window (cudaMalloc)
    size_t win_bytes = 0, heap_used = 0;
    PeerDev pd;
    int* h_err = nullptr;               // watchdog latch of the in-kernel waits (mapped pinned; see PeerDev)
    unsigned long long seq[5] = {0, 0, 0, 0, 0};   // channels 0..3 + halo
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

// ------------------------------------------------------------------ rendezvous ----
static int32_t rv_open(b2k_ctx* ctx, Rendezvous* rv, const void* uid128, int rank, int nranks) {
    unsigned long long h = 1469598103934665603ull;                 // FNV-1a of the unique id
    const unsigned char* u = (const unsigned char*)uid128;
    for (int i = 0; i < 128; ++i) { h ^= u[i]; h *= 1099511628211ull; }
    snprintf(rv->name, sizeof(rv->name), "/b2k_%016llx_%d", h, nranks);
    const int fd = shm_open(rv->name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return b2k_fail(ctx, B2K_ENCCL, "rendezvous: shm_open(%s) failed", rv->name);
    if (ftruncate(fd, sizeof(RvShm)) != 0) {                       // new pages are zero-filled
        close(fd);
        return b2k_fail(ctx, B2K_ENCCL, "rendezvous: ftruncate failed");
    }
    void* p = mmap(nullptr, sizeof(RvShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return b2k_fail(ctx, B2K_ENCCL, "rendezvous: mmap failed");
    rv->shm = (RvShm*)p;
    rv->rank = rank;
    rv->nranks = nranks;
    rv->round = 0;
    return B2K_OK;
}

static void rv_close(Rendezvous* rv) {
    if (!rv->shm) return;
    munmap(rv->shm, sizeof(RvShm));
    rv->shm = nullptr;
}

// all[p*bytes ..] = rank p's blob.  A rank cannot be more than one round ahead of a peer (the next round
// needs that peer's arrival), so two data buffers are enough.
static int32_t rv_allgather(b2k_ctx* ctx, Rendezvous* rv, const void* mine, size_t bytes, void* all) {
    if (!rv->shm) return b2k_fail(ctx, B2K_ENCCL, "rendezvous not open");
    if (bytes > RV_BLOB) return b2k_fail(ctx, B2K_EINVAL, "rendezvous: blob too large");
    const unsigned long long r = ++rv->round;
    const int buf = (int)(r & 1ull);
    memcpy(rv->shm->data[buf][rv->rank], mine, bytes);
    __atomic_fetch_add(&rv->shm->arrive, 1ull, __ATOMIC_SEQ_CST);
    const unsigned long long target = r * (unsigned long long)rv->nranks;
    const time_t t0 = time(nullptr);
    while (__atomic_load_n(&rv->shm->arrive, __ATOMIC_SEQ_CST) < target) {
        usleep(50);
        if (time(nullptr) - t0 > 120)
            return b2k_fail(ctx, B2K_ENCCL, "rendezvous: timed out waiting for %d ranks (round %llu)", rv->nranks, r);
    }
    for (int p = 0; p < rv->nranks; ++p) memcpy((char*)all + (size_t)p * bytes, rv->shm->data[buf][p], bytes);
    return B2K_OK;
}

// ------------------------------------------------------------------ peer window ----
// generic small-vector all-reduce (channel 3): publish, wait, sum in rank order (same bits on every rank)
__global__ void __launch_bounds__(256)
k_peer_allreduce(const __grid_constant__ PeerDev pd, double* __restrict__ inout, int count,
                 unsigned long long seq) {
    for (int idx = threadIdx.x; idx < count * pd.nranks; idx += blockDim.x) {
        const int p = idx / count, j = idx - p * count;
        peer_slot(pd, p, PEER_CH_GEN, seq, pd.rank)[j] = inout[j];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pd.nranks)
        st_release_sys_u64(peer_flag(pd, threadIdx.x, PEER_CH_GEN, seq, pd.rank), seq);
    peer_wait(pd, PEER_CH_GEN, seq, threadIdx.x);
    __syncthreads();
    for (int j = threadIdx.x; j < count; j += blockDim.x) inout[j] = peer_sum1(pd, PEER_CH_GEN, seq, j);
}

static size_t window_bytes() {
    const char* e = getenv("B2K_PEER_WINDOW_MB");
    const double mb = e ? atof(e) : 48.0;
    return PEER_OFF_HEAP + (size_t)(mb * 1048576.0);
}

static void peer_setup(b2k_ctx* ctx, B2kNccl* n) {
    // NOTE: every rank runs the same rendezvous rounds below, whatever its local outcome.
    n->peer_ok = false;
    const char* off = getenv("B2K_PEER");
    if ((off && off[0] == '0') || ctx->nranks > PEER_MAXR || !n->rv.shm) return;   // same decision on every rank
    const int R = ctx->nranks;
    n->win_bytes = window_bytes();
    int good = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&n->win_local, n->win_bytes) != cudaSuccess) { cudaGetLastError(); n->win_local = nullptr; good = 0; }
    if (good && cudaMemset(n->win_local, 0, n->win_bytes) != cudaSuccess) { cudaGetLastError(); good = 0; }
    if (good && cudaIpcGetMemHandle(&mine, n->win_local) != cudaSuccess) { cudaGetLastError(); good = 0; }
    cudaDeviceSynchronize();
    struct Blob { cudaIpcMemHandle_t h; int good; int pid; } me, all[PEER_MAXR];
    memset(&me, 0, sizeof(me));
    me.h = mine; me.good = good; me.pid = (int)getpid();
    if (rv_allgather(ctx, &n->rv, &me, sizeof(me), all) != B2K_OK) return;
    for (int p = 0; p < R; ++p) good = good && all[p].good;
    memset(&n->pd, 0, sizeof(n->pd));
    for (int p = 0; p < R && good; ++p) {
        if (p == ctx->rank) { n->pd.win[p] = n->win_local; continue; }
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[p].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            good = 0;
            break;
        }
        n->pd.win[p] = (char*)q;
    }
    // every rank must agree (and have finished mapping) before the first use
    int mygood = good, allgood[PEER_MAXR];
    if (rv_allgather(ctx, &n->rv, &mygood, sizeof(int), allgood) != B2K_OK) return;
    for (int p = 0; p < R; ++p) good = good && allgood[p];
    n->pd.rank = ctx->rank;
    n->pd.nranks = R;
    n->peer_ok = good != 0;
    if (n->peer_ok) {
        // watchdog latch: one int the kernels can write and the host can read without a copy.  Without it (the
        // allocation failed) the waits are the plain unbounded loops.
        const char* e = getenv("B2K_PEER_TIMEOUT_S");
        const double secs = e ? atof(e) : 120.0;
        void* dp = nullptr;
        if (cudaHostAlloc((void**)&n->h_err, sizeof(int), cudaHostAllocMapped) == cudaSuccess &&
            cudaHostGetDevicePointer(&dp, n->h_err, 0) == cudaSuccess) {
            *n->h_err = 0;
            n->pd.err = (int*)dp;
            n->pd.timeout_ns = secs > 0.0 ? (unsigned long long)(secs * 1e9) : 0ull;
        } else {
            cudaGetLastError();
            if (n->h_err) cudaFreeHost(n->h_err);
            n->h_err = nullptr;
            n->pd.err = nullptr;
            n->pd.timeout_ns = 0;
        }
    }
    n->heap_used = 0;
    for (int c = 0; c < 5; ++c) n->seq[c] = 0;
    if (!n->peer_ok && n->win_local) {
        for (int p = 0; p < R; ++p)
            if (p != ctx->rank && n->pd.win[p]) cudaIpcCloseMemHandle(n->pd.win[p]);
        cudaFree(n->win_local);
        n->win_local = nullptr;
    }
}

bool b2k_peer_ok(const b2k_ctx* ctx) { return ctx->nccl && ctx->nccl->peer_ok; }

int32_t b2k_stream_sync(b2k_ctx* ctx) {
    B2K_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const B2kNccl* n = ctx->nccl;
    if (n && n->h_err && *(volatile int*)n->h_err != 0)
        return b2k_fail(ctx, B2K_ENCCL,
                        "peer window: an in-kernel wait for another rank's flag timed out (B2K_PEER_TIMEOUT_S, "
                        "default 120 s) - a rank of the job has died or left the SPMD call order; the results of "
                        "this rank are invalid and the context must be destroyed");
    return B2K_OK;
}
bool b2k_has_nccl(const b2k_ctx* ctx) { return ctx->nccl && ctx->nccl->comm != nullptr; }
const PeerDev* b2k_peer_dev(const b2k_ctx* ctx) { return &ctx->nccl->pd; }
char* b2k_peer_local(const b2k_ctx* ctx) { return ctx->nccl->win_local; }
unsigned long long b2k_peer_next_seq(b2k_ctx* ctx, int channel) { return ++ctx->nccl->seq[channel]; }

size_t b2k_peer_heap_alloc(b2k_ctx* ctx, size_t bytes) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->peer_ok) return SIZE_MAX;
    const size_t off = PEER_OFF_HEAP + ((n->heap_used + 255) & ~(size_t)255);
    if (off + bytes > n->win_bytes) return SIZE_MAX;
    n->heap_used = off + bytes - PEER_OFF_HEAP;
    return off;
}

int32_t b2k_host_allgather(b2k_ctx* ctx, const void* mine, size_t bytes, void* all) {
    if (!ctx->nccl) return b2k_fail(ctx, B2K_ENCCL, "host allgather on a context without a rendezvous");
    return rv_allgather(ctx, &ctx->nccl->rv, mine, bytes, all);
}

int32_t b2k_peer_allreduce(b2k_ctx* ctx, double* dptr, int32_t count) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->peer_ok || count > PEER_SLOT) return b2k_fail(ctx, B2K_ENCCL, "peer all-reduce unavailable");
    const unsigned long long seq = ++n->seq[PEER_CH_GEN];
    k_peer_allreduce<<<1, 256, 0, ctx->stream>>>(n->pd, dptr, count, seq);
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
    if (g_comm_cache && !g_comm_in_use && g_comm_cache->h_err && *(volatile int*)g_comm_cache->h_err != 0)
        return b2k_fail(ctx, B2K_ENCCL, "the peer window of this process is out of step with its peers (an "
                                        "in-kernel wait timed out in an earlier context); restart the job");
    if (g_comm_cache && !g_comm_in_use && g_comm_rank == ctx->rank && g_comm_nranks == ctx->nranks &&
        g_comm_device == ctx->device) {
        ctx->nccl = g_comm_cache;
        g_comm_in_use = true;
        return B2K_OK;
    }
    if (!uid) return b2k_fail(ctx, B2K_EINVAL, "ctx_create_dist: nccl_uid is NULL");
    B2kNccl* n = new B2kNccl();
    ctx->nccl = n;
    // B2K_NO_NCCL=1: peer windows only (the uid is then just a 128-byte job-unique token).  That is the mode
    // in which several ranks may share one GPU.
    const char* nonccl = getenv("B2K_NO_NCCL");
    const bool want_nccl = !(nonccl && nonccl[0] == '1');
    int32_t rc = rv_open(ctx, &n->rv, uid, ctx->rank, ctx->nranks);
    if (rc != B2K_OK) { ctx->nccl = nullptr; delete n; return rc; }
    {   // one round proves that every rank has mapped the segment; its name can then go away (the mapping
        // stays valid), so nothing is left behind in /dev/shm even though the object is cached for the process
        int pids[PEER_MAXR], me = (int)getpid();
        rc = rv_allgather(ctx, &n->rv, &me, sizeof(int), pids);
        shm_unlink(n->rv.name);
        if (rc != B2K_OK) { rv_close(&n->rv); ctx->nccl = nullptr; delete n; return rc; }
    }
    if (want_nccl) {
        rc = load_nccl(ctx, n);
        if (rc != B2K_OK) { rv_close(&n->rv); ctx->nccl = nullptr; delete n; return rc; }
        ncclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        NCCL_CK(ctx, n, n->CommInitRank(&n->comm, ctx->nranks, id, ctx->rank));
    }
    peer_setup(ctx, n);
    if (!n->peer_ok && !n->comm) {
        rv_close(&n->rv);
        ctx->nccl = nullptr;
        delete n;
        return b2k_fail(ctx, B2K_ENCCL, "no transport: NCCL disabled (B2K_NO_NCCL=1) and the NVLink peer window "
                                        "could not be mapped on every rank");
    }
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
    ctx->nccl->heap_used = 0;            // operators (the heap's only users) die with their context
    if (ctx->nccl == g_comm_cache) {     // stays alive for the next context of this process
        g_comm_in_use = false;
        ctx->nccl = nullptr;
        return;
    }
    if (ctx->nccl->peer_ok)
        for (int p = 0; p < ctx->nranks; ++p)
            if (p != ctx->rank) cudaIpcCloseMemHandle(ctx->nccl->pd.win[p]);
    if (ctx->nccl->win_local) cudaFree(ctx->nccl->win_local);
    if (ctx->nccl->h_err) cudaFreeHost(ctx->nccl->h_err);
    rv_close(&ctx->nccl->rv);
    if (ctx->nccl->comm) ctx->nccl->CommDestroy(ctx->nccl->comm);
    delete ctx->nccl;
    ctx->nccl = nullptr;
}

int32_t b2k_nccl_allreduce_f64(b2k_ctx* ctx, double* dptr, int32_t count) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->comm) return b2k_fail(ctx, B2K_ENCCL, "allreduce on a context without a communicator");
    NCCL_CK(ctx, n, n->AllReduce(dptr, dptr, (size_t)count, ncclFloat64_, ncclSum_, n->comm, ctx->stream));
    return B2K_OK;
}

int32_t b2k_nccl_allgather(b2k_ctx* ctx, const void* sendbuf, void* recvbuf, size_t bytes) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->comm) return b2k_fail(ctx, B2K_ENCCL, "allgather on a context without a communicator");
    NCCL_CK(ctx, n, n->AllGather(sendbuf, recvbuf, bytes, ncclChar_, n->comm, ctx->stream));
    return B2K_OK;
}

int32_t b2k_nccl_halo_exchange(b2k_ctx* ctx, int up, int dn, const void* send_up, size_t send_up_bytes,
                               void* recv_dn, size_t recv_dn_bytes, const void* send_dn,
                               size_t send_dn_bytes, void* recv_up, size_t recv_up_bytes) {
    B2kNccl* n = ctx->nccl;
    if (!n || !n->comm) return b2k_fail(ctx, B2K_ENCCL, "halo exchange on a context without a communicator");
    NCCL_CK(ctx, n, n->GroupStart());
    if (up >= 0 && send_up_bytes) NCCL_CK(ctx, n, n->Send(send_up, send_up_bytes, ncclChar_, up, n->comm, ctx->stream));
    if (dn >= 0 && recv_dn_bytes) NCCL_CK(ctx, n, n->Recv(recv_dn, recv_dn_bytes, ncclChar_, dn, n->comm, ctx->stream));
    if (dn >= 0 && send_dn_bytes) NCCL_CK(ctx, n, n->Send(send_dn, send_dn_bytes, ncclChar_, dn, n->comm, ctx->stream));
    if (up >= 0 && recv_up_bytes) NCCL_CK(ctx, n, n->Recv(recv_up, recv_up_bytes, ncclChar_, up, n->comm, ctx->stream));
    NCCL_CK(ctx, n, n->GroupEnd());
    return B2K_OK;
}
