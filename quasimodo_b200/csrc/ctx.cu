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

extern "C" {

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
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return QM_ECUDA; }
    *out = c;
    return QM_OK;
}

void qm_ctx_destroy(qm_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->scratch) if (s.ptr) cudaFree(s.ptr);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *qm_last_error(const qm_ctx *ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
int qm_device_sm_count(const qm_ctx *ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"
