// comm.cu -- the one collective of the path: per-GPU int32 count tensors summed over NVLink (north_star: "per-GPU int32
// count tensors are merged with one NCCL allreduce"), plus the 128-byte broadcast of the insert-size
// model so that every rank pairs its reads against the same mem_pestat result.
//
// NCCL is bound at RUN time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, loads on hosts without
// NCCL (single-GPU use, the CPU test box), and inside a torch process resolves to the very libnccl torch already loaded.
// Two ways to get a communicator: one process per GPU (qm_comm_unique_id on rank 0, the 128 bytes carried to the other
// ranks by whatever the host has -- torch.distributed, MPI, a file -- then qm_comm_init_rank), or one process driving N GPUs
// (qm_comm_init_all, what `qm_driver sample --gpus` uses).  qm_counts_allreduce_nccl takes a caller-owned ncclComm_t.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "common.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclInt32 = 2 };
enum { ncclSum = 0 };

struct Nccl {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string why;
};

Nccl g_nccl;

Nccl *nccl()
{
    static std::once_flag once;
    std::call_once(once, [] {
        // QM_NCCL_LIB=/path/to/libnccl.so.2 names the library when it is not on the loader's path (and nothing else is tried)
        const char *env = getenv("QM_NCCL_LIB");
        const char *names[] = {env && *env ? env : "libnccl.so.2", env && *env ? nullptr : "libnccl.so"};
        for (const char *nm : names) { if (nm) g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
        if (!g_nccl.h) {
            const char *err = dlerror();               // one call: dlerror() clears the message it returns
            g_nccl.why = std::string("dlopen(") + names[0] + "): " + (err ? err : "not found");
            return;
        }
        auto sym = [&](const char *s) { void *p = dlsym(g_nccl.h, s); if (!p && g_nccl.why.empty()) g_nccl.why = std::string("libnccl lacks ") + s; return p; };
        g_nccl.GetUniqueId = (int (*)(ncclUniqueId *))sym("ncclGetUniqueId");
        g_nccl.CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))sym("ncclCommInitRank");
        g_nccl.CommInitAll = (int (*)(ncclComm_t *, int, const int *))sym("ncclCommInitAll");
        g_nccl.CommDestroy = (int (*)(ncclComm_t))sym("ncclCommDestroy");
        g_nccl.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclAllReduce");
        g_nccl.Broadcast = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))sym("ncclBroadcast");
        g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))sym("ncclAllGather");
        g_nccl.GetVersion = (int (*)(int *))sym("ncclGetVersion");
        g_nccl.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    });
    return (g_nccl.h && g_nccl.why.empty()) ? &g_nccl : nullptr;
}

// why nccl() answered null (call it after nccl())
const char *nccl_why()
{
    static std::string msg;
    msg = "NCCL is not available: " + (g_nccl.why.empty() ? std::string("libnccl.so.2 could not be loaded") : g_nccl.why);
    return msg.c_str();
}

}  // namespace

struct qm_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    bool owned = true;
};

#define QM_NCCL(ctx, call)                                                                                     \
    do {                                                                                                       \
        const int r__ = (call);                                                                                \
        if (r__ != ncclSuccess)                                                                                \
            return qm_fail((ctx), QM_ECUDA, "%s:%d %s -> NCCL error %d (%s)", __FILE__, __LINE__, #call, r__,  \
                           N->GetErrorString ? N->GetErrorString(r__) : "?");                                  \
    } while (0)

extern "C" {

int qm_comm_available(void) { return nccl() != nullptr; }

int qm_comm_unique_id(uint8_t id[128])
{
    Nccl *N = nccl();
    if (!N || !id) return N ? QM_EINVAL : QM_ENODEV;
    ncclUniqueId u;
    if (N->GetUniqueId(&u) != ncclSuccess) return QM_ECUDA;
    memcpy(id, u.internal, 128);
    return QM_OK;
}

int qm_comm_init_rank(qm_ctx *ctx, int n_ranks, int rank, const uint8_t id[128], qm_comm **out)
{
    if (!ctx || !id || !out || n_ranks < 1 || rank < 0 || rank >= n_ranks) return QM_EINVAL;
    *out = nullptr;
    Nccl *N = nccl();
    if (!N) return qm_fail(ctx, QM_ENODEV, "%s", nccl_why());
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    qm_comm *c = new qm_comm;
    c->rank = rank; c->size = n_ranks;
    const int r = N->CommInitRank(&c->comm, n_ranks, u, rank);
    if (r != ncclSuccess) { delete c; return qm_fail(ctx, QM_ECUDA, "ncclCommInitRank(%d of %d) -> %d (%s)", rank, n_ranks, r, N->GetErrorString(r)); }
    *out = c;
    return QM_OK;
}

// one process, n GPUs: comms[i] belongs to ctxs[i]'s device (rank i)
int qm_comm_init_all(int n, qm_ctx *const *ctxs, qm_comm **comms)
{
    if (n < 1 || !ctxs || !comms) return QM_EINVAL;
    for (int i = 0; i < n; ++i) { if (!ctxs[i]) return QM_EINVAL; comms[i] = nullptr; }
    Nccl *N = nccl();
    if (!N) return qm_fail(ctxs[0], QM_ENODEV, "%s", nccl_why());
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    std::vector<ncclComm_t> cs(n, nullptr);
    const int r = N->CommInitAll(cs.data(), n, devs.data());
    if (r != ncclSuccess) return qm_fail(ctxs[0], QM_ECUDA, "ncclCommInitAll(%d devices) -> %d (%s)", n, r, N->GetErrorString(r));
    for (int i = 0; i < n; ++i) { comms[i] = new qm_comm; comms[i]->comm = cs[i]; comms[i]->rank = i; comms[i]->size = n; }
    return QM_OK;
}

void qm_comm_destroy(qm_comm *c)
{
    if (!c) return;
    Nccl *N = nccl();
    if (N && c->comm && c->owned) N->CommDestroy(c->comm);
    delete c;
}

int qm_comm_rank(const qm_comm *c) { return c ? c->rank : -1; }
int qm_comm_size(const qm_comm *c) { return c ? c->size : 0; }

// in-place sum of n int32 over the communicator's ranks (integer sum: order independent, bit-exact for any number of GPUs);
// asynchronous on `stream`.  The variant below takes a caller-owned ncclComm_t (SURVEY.md 8b's signature).
int qm_counts_allreduce_nccl(qm_ctx *ctx, void *nccl_comm, int32_t *d_counts, int64_t n, void *stream)
{
    if (!ctx || !nccl_comm || n < 0 || (n > 0 && !d_counts)) return QM_EINVAL;
    Nccl *N = nccl();
    if (!N) return qm_fail(ctx, QM_ENODEV, "%s", nccl_why());
    if (n == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int sp = qm_prof_begin(ctx, QM_ST_OTHER, st);
    QM_NCCL(ctx, N->AllReduce(d_counts, d_counts, (size_t)n, ncclInt32, ncclSum, (ncclComm_t)nccl_comm, st));
    qm_prof_end(ctx, QM_ST_OTHER, sp, st, 1);
    return QM_OK;
}

int qm_counts_allreduce(qm_ctx *ctx, qm_comm *comm, int32_t *d_counts, int64_t n, void *stream)
{
    if (!comm) return QM_EINVAL;
    if (comm->size == 1) return QM_OK;
    return qm_counts_allreduce_nccl(ctx, comm->comm, d_counts, n, stream);
}

// every rank's `bytes` bytes at d_send, concatenated in rank order into d_recv (size * bytes) on every rank; asynchronous
int qm_comm_allgather(qm_ctx *ctx, qm_comm *comm, const void *d_send, void *d_recv, size_t bytes, void *stream)
{
    if (!ctx || !comm || (bytes > 0 && (!d_send || !d_recv))) return QM_EINVAL;
    if (bytes == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (comm->size == 1) {
        if (d_send != d_recv) QM_CUDA(ctx, cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, st));
        return QM_OK;
    }
    Nccl *N = nccl();
    if (!N) return qm_fail(ctx, QM_ENODEV, "%s", nccl_why());
    QM_NCCL(ctx, N->AllGather(d_send, d_recv, bytes, ncclInt8, comm->comm, st));
    return QM_OK;
}

// the insert-size model of rank `root` to every rank (4 x qm_pestat = 128 bytes); synchronous
int qm_pestat_bcast(qm_ctx *ctx, qm_comm *comm, qm_pestat pes[4], int root, void *stream)
{
    if (!ctx || !comm || !pes || root < 0 || root >= comm->size) return QM_EINVAL;
    if (comm->size == 1) return QM_OK;
    Nccl *N = nccl();
    if (!N) return qm_fail(ctx, QM_ENODEV, "%s", nccl_why());
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 20, 256, &p);
    if (rc) return rc;
    if (comm->rank == root) QM_CUDA(ctx, cudaMemcpyAsync(p, pes, 4 * sizeof(qm_pestat), cudaMemcpyHostToDevice, st));
    QM_NCCL(ctx, N->Broadcast(p, p, 4 * sizeof(qm_pestat), ncclInt8, root, comm->comm, st));
    QM_CUDA(ctx, cudaMemcpyAsync(pes, p, 4 * sizeof(qm_pestat), cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    return QM_OK;
}

}  // extern "C"
