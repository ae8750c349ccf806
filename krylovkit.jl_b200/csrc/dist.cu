// dist.cu — NCCL plumbing for row-sharded contexts (one process per GPU).
// NCCL is resolved at run time with dlopen so that (a) the library has no link-time NCCL
// dependency and loads on machines without it, and (b) inside a PyTorch process we bind to
// the very libnccl.so.2 torch already loaded instead of a second copy.
// Collectives per step: AllReduce of <= k+1 doubles (latency-bound) and a
// nearest-neighbour halo exchange for the SpMV.
#include "common.cuh"
#include <dlfcn.h>

// minimal NCCL ABI (stable since 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess_ = 0 };
enum { ncclInt8_ = 0, ncclChar_ = 0, ncclFloat64_ = 8 };
enum { ncclSum_ = 0 };

struct B2kNccl {
    void* lib = nullptr;
    ncclComm_t comm = nullptr;
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

int32_t b2k_nccl_init(b2k_ctx* ctx, const void* uid) {
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
    return B2K_OK;
}

void b2k_nccl_destroy(b2k_ctx* ctx) {
    if (!ctx->nccl) return;
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
