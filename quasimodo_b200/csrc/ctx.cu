// ctx.cu -- context, error reporting, scratch memory.  No CPU fallback: creation fails without a device.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

int qm_fail(qm_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int qm_scratch_reserve(qm_ctx *ctx, int which, size_t bytes, void **out)
{
    qm_scratch &s = ctx->scratch[which];
    if (bytes > s.cap) {
        if (s.ptr) QM_CUDA(ctx, cudaFree(s.ptr));
        s.ptr = nullptr; s.cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        QM_CUDA(ctx, cudaMalloc(&s.ptr, want));
        s.cap = want;
    }
    *out = s.ptr;
    return QM_OK;
}

static cudaEvent_t prof_event(qm_ctx *ctx)
{
    cudaEvent_t e = nullptr;
    if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

int qm_prof_begin(qm_ctx *ctx, int stage, cudaStream_t st)
{
    if (!ctx->prof_on) return -1;
    qm_prof_span sp;
    sp.stage = stage; sp.e0 = prof_event(ctx); sp.e1 = prof_event(ctx);
    cudaEventRecord(sp.e0, st);
    ctx->prof_spans.push_back(sp);
    return (int)ctx->prof_spans.size() - 1;
}

void qm_prof_end(qm_ctx *ctx, int stage, int span, cudaStream_t st, int launches)
{
    ctx->prof_launches[stage] += launches;          // launch counts are kept even when timing is off
    if (span >= 0) cudaEventRecord(ctx->prof_spans[span].e1, st);
}

extern "C" {

int qm_profile_enable(qm_ctx *ctx, int on)
{
    if (!ctx) return QM_EINVAL;
    ctx->prof_on = on != 0;
    return QM_OK;
}

// synchronises the device, adds every finished span to the per-stage totals, returns and clears them
int qm_profile_collect(qm_ctx *ctx, double *ms_out, int64_t *launches_out)
{
    if (!ctx) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    QM_CUDA(ctx, cudaDeviceSynchronize());
    for (auto &sp : ctx->prof_spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) ctx->prof_ms[sp.stage] += ms;
        ctx->prof_pool.push_back(sp.e0); ctx->prof_pool.push_back(sp.e1);
    }
    ctx->prof_spans.clear();
    for (int i = 0; i < QM_ST_N; ++i) {
        if (ms_out) ms_out[i] = ctx->prof_ms[i];
        if (launches_out) launches_out[i] = ctx->prof_launches[i];
        ctx->prof_ms[i] = 0; ctx->prof_launches[i] = 0;
    }
    return QM_OK;
}

void qm_opt_default(qm_opt *o)
{
    memset(o, 0, sizeof *o);
    o->a = 1; o->b = 4;
    o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
    o->w = 100; o->zdrop = 100;
    o->pen_clip5 = o->pen_clip3 = 5;
    o->min_seed_len = 31;              // rules/bwa.smk:15 `-k 31`
    o->max_occ = 500;
    o->T = 30;
    o->pen_unpaired = 17;
    o->max_ins = 10000;
    o->max_chain_gap = 10000;
    o->mapq_coef_len = 50;
    o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->mask_level_redun = 0.95f;
    o->min_chain_weight = 0;
}

const char *qm_version(void) { return "quasimodo_b200 0.1 (sm_100a)"; }

int qm_ctx_create(int device, qm_ctx **out)
{
    if (!out) return QM_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return QM_ENODEV;
    if (cudaSetDevice(device) != cudaSuccess) return QM_ENODEV;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return QM_ENODEV;
    if (prop.major < 10) return QM_ENODEV;           // built for sm_100a only
    qm_ctx *c = new qm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return QM_ECUDA; }
    for (int i = 0; i < 12; ++i) {
        if (cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming) != cudaSuccess) { qm_ctx_destroy(c); return QM_ECUDA; }
    }
    if (cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess) { qm_ctx_destroy(c); return QM_ECUDA; }
    if (cudaMallocHost(&c->h_pinned, 8192) != cudaSuccess) { qm_ctx_destroy(c); return QM_ENOMEM; }
    *out = c;
    return QM_OK;
}

void qm_ctx_destroy(qm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->scratch) if (s.ptr) cudaFree(s.ptr);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 12; ++i) {
        if (ctx->side[i]) cudaStreamDestroy(ctx->side[i]);
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (auto &sp : ctx->prof_spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    delete ctx;
}

const char *qm_last_error(const qm_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
int qm_device_sm_count(const qm_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
