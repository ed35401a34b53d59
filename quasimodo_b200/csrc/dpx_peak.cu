// dpx_peak.cu -- measures the whole-GPU issue rate of the DPX instructions the extension kernel is
// built on (BASELINE.md section 2 asks the builder to measure it; SURVEY.md 8d defines the GCUPS
// roofline as R_dpx * lanes / 9).  Register-resident, 8 independent dependent-chains per thread.
#include "common.cuh"

namespace {
template <int KIND>
__global__ void __launch_bounds__(256) dpx_kernel(int iters, int seed, int *sink)
{
    int x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = seed + threadIdx.x * 8 + k;
    const int b = seed | 1, c = seed >> 3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (KIND == 0) x[k] = __viaddmax_s32_relu(x[k], -b, c + k);
                else if (KIND == 1) x[k] = (int)__viaddmax_s16x2_relu((unsigned)x[k], (unsigned)b, (unsigned)(c + k));
                else x[k] = __vimax3_s32(x[k] - 1, b, c + k);
            }
        }
    }
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc ^= x[k];
    if (acc == 0x7fffffff) sink[0] = acc;     // keeps the chains alive
}
}  // namespace

extern "C" int qm_dpx_peak_sync(qm_ctx *ctx, int kind, int iters, double *out_gops, double *out_ms)
{
    if (!ctx || !out_gops || iters <= 0 || kind < 0 || kind > 2) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 2, 256, &p);
    if (rc) return rc;
    const int blocks = ctx->sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    QM_CUDA(ctx, cudaEventCreate(&e0));
    QM_CUDA(ctx, cudaEventCreate(&e1));
    cudaStream_t st = ctx->own_stream;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        QM_CUDA(ctx, cudaEventRecord(e0, st));
        if (kind == 0) dpx_kernel<0><<<blocks, threads, 0, st>>>(iters, 12345, (int *)p);
        else if (kind == 1) dpx_kernel<1><<<blocks, threads, 0, st>>>(iters, 12345, (int *)p);
        else dpx_kernel<2><<<blocks, threads, 0, st>>>(iters, 12345, (int *)p);
        QM_CUDA(ctx, cudaEventRecord(e1, st));
        QM_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        QM_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    const double ops = (double)blocks * threads * (double)iters * 64.0;    // lane-instructions
    *out_gops = ops / (best * 1e-3) / 1e9;
    if (out_ms) *out_ms = best;
    return QM_OK;
}
