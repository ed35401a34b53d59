// extend.cu -- batched banded affine-gap extension, bit-exact with bwa's ksw_extend2
// (bwa 0.7.17 ksw.c, called from bwamem.c:mem_chain2aln; reference call site rules/bwa.smk:15;
// semantics restated in SURVEY.md Appendix A.3).
//
// Formulation ("B" of SURVEY.md A.3): one warp per task, one DP row (target base) per step, the
// query columns striped across the lanes in blocks of C consecutive columns.  Each lane keeps its
// slice of the reference's in-place eh[] array in registers, so columns a row does not visit keep
// their stale values exactly like the CPU loop.  Because gaps open from M only, the F recurrence
// inside a row is a max-plus prefix scan: lane-local chain + one 5-step warp scan.  Row controls
// (band, (m,mj), z-drop, to-end score, zero-span trimming) are warp-uniform, computed from three
// REDUX reductions.  Cell arithmetic uses the Blackwell DPX instructions
// (__viaddmax_s32_relu / __vimax3_s32 / __viaddmax_s32); no tensor cores: this is not a contraction.
#include <stdlib.h>
#include "pipeline.cuh"
#include "ext_warp.cuh"

namespace {

constexpr int kNumClasses = kExtClasses;
constexpr int kThreadPerTaskMin = 16384;       // tasks in a class below which the warp-per-task kernel is used (a thread of the
                                               // packed kernel holds two tasks: fewer would leave most SMs without a warp)

// ---- public tasks -> internal tasks, binned by striping class (device side, no host sync) ----
__global__ void ext_classify_kernel(const qm_ext_task *__restrict__ tasks, const uint8_t *__restrict__ seq, int64_t n,
                                    ExtTaskI *__restrict__ itasks, int *__restrict__ lists, int *__restrict__ counts,
                                    qm_ext_result *__restrict__ out, int *__restrict__ err)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const qm_ext_task t = tasks[i];
    if (t.qlen < 0 || t.tlen < 0 || t.qlen > QM_EXT_MAX_QLEN || t.h0 <= 0 || t.w < 0) {
        qm_ext_result r = {};
        r.score = INT32_MIN;
        out[i] = r;
        atomicExch(err, 1);
        return;
    }
    ExtTaskI it;
    it.q = seq + t.q_off; it.t = seq + t.t_off; it.t0 = 0; it.qstep = 1; it.tstep = 1;
    it.qlen = t.qlen; it.tlen = t.tlen; it.h0 = t.h0; it.w = t.w; it.end_bonus = t.end_bonus;
    it.flags = t.flags & (QM_EXT_BAND_RETRY | QM_EXT_PREV_H0);
    it.pad[0] = it.pad[1] = 0;
    itasks[i] = it;
    const int c = qm_ext_class(t.qlen);
    const int slot = atomicAdd(&counts[c], 1);
    lists[(int64_t)c * n + slot] = (int)i;
}

template <int C>
__global__ void __launch_bounds__(128)
ext_kernel(ExtParams P, IndexView V, const ExtTaskI *__restrict__ tasks, const int *__restrict__ list,
           const int *__restrict__ count, int *__restrict__ cursor, qm_ext_result *__restrict__ out)
{
    const int lane = qm_lane();
    const int n = *count;
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(cursor, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= n) break;
        const int tid = list[idx];
        const ExtTaskI t = tasks[tid];
        SeqFetch S;
        S.q = t.q; S.t = t.t; S.t0 = t.t0; S.qstep = t.qstep; S.tstep = t.tstep;
        S.indirect = (t.flags & QM_EXTI_INDIRECT) != 0; S.V = &V;
        ExtState r;
        int w_used = t.w, cells = 0, prev = (t.flags & QM_EXT_PREV_H0) ? t.h0 : -1;
        const int tries = (t.flags & QM_EXT_BAND_RETRY) ? 2 : 1;
        for (int a = 0; a < tries; ++a) {
            w_used = t.w << a;
            r = ext_run<C>(P, S, t.qlen, t.tlen, t.h0, w_used, t.end_bonus, lane);
            cells += r.cells;
            if (r.score == prev || r.max_off < (w_used >> 1) + (w_used >> 2)) break;
            prev = r.score;
        }
        if (lane == 0) {
            qm_ext_result o;
            o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore;
            o.max_off = r.max_off; o.w_used = w_used; o.cells = cells;
            out[tid] = o;
        }
    }
}

template <int C>
void launch_class(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks, const int *d_list,
                  const int *d_counts, int *d_cursors, const int *h_counts, qm_ext_result *d_out,
                  cudaStream_t st)
{
    int64_t warps = (int64_t)ctx->sm_count * 64;          // persistent: up to 16 blocks of 4 warps per SM
    if (h_counts) {
        if (h_counts[cls] == 0) return;
        if (h_counts[cls] < warps) warps = h_counts[cls];
    }
    unsigned blocks = (unsigned)((warps + 3) / 4);
    if (blocks > (unsigned)ctx->sm_count) blocks = ((blocks + ctx->sm_count - 1) / ctx->sm_count) * ctx->sm_count;
    ext_kernel<C><<<blocks, 128, 0, st>>>(P, V, d_tasks, d_list, d_counts + cls, d_cursors + cls, d_out);
}

}  // namespace

int qm_ext_launch_classes(qm_ctx *ctx, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks,
                          const int *d_lists, int64_t list_stride, const int *d_counts, int *d_cursors,
                          const int *h_counts, qm_ext_result *d_out, int *d_fb_lists, int *d_fb_ctr, cudaStream_t st,
                          bool scores_fit_bytes, const int64_t *h_list_off)
{
    // A class with many tasks goes to the thread-per-task kernel (extend2.cu: high throughput, but one task is a
    // long serial chain, ~0.3 ms); a class with few tasks (the tail rounds of mem_chain2aln, where only reads with
    // many chains are still active) goes to the warp-per-task kernel (low latency).  Class 9 (qlen > 256) always
    // does.  Without host-side counts (public qm_extend_batch) classes 0..8 use the thread-per-task kernel.
    // The classes of one round are independent: each runs on its own side stream, forked from / joined to `st`.
    QM_CUDA(ctx, cudaEventRecord(ctx->ev_fork, st));
    for (int c = 0; c < kExtClasses; ++c) {
        if (h_counts && h_counts[c] == 0) continue;
        cudaStream_t sc = ctx->side[c];
        QM_CUDA(ctx, cudaStreamWaitEvent(sc, ctx->ev_fork, 0));
        const int *lst = h_list_off ? d_lists + h_list_off[c] : d_lists + c * list_stride;
        int *fbl = d_fb_lists ? d_fb_lists + c * list_stride : nullptr;
        static const int tpt_min = getenv("QM_TPT_MIN") ? atoi(getenv("QM_TPT_MIN")) : kThreadPerTaskMin;      // tuning knob
        const bool big = c < 9 && (!h_counts || h_counts[c] >= tpt_min);
        // The packed two-tasks-per-thread kernel (extend3.cu) takes every class it can hold; QM_EXT3=0 keeps the scalar
        // thread-per-task kernel (extend2.cu) for A/B measurements.
        static const bool ext3 = !(getenv("QM_EXT3") && atoi(getenv("QM_EXT3")) == 0);
        if (big && d_fb_lists && d_fb_ctr && ext3 && qm_ext3_scores_ok(P)) {
            const int hc = h_counts ? h_counts[c] : -1;
            int rc = qm_ext3_launch_class(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, hc, d_out, fbl, d_fb_ctr, sc);
            if (rc) return rc;
            // what the packed arithmetic cannot hold (scores above 255) comes back through the fallback list; the caller
            // that knows the round's maximum score (scores_fit_bytes) knows the list stays empty
            if (!scores_fit_bytes) {
                switch (c) {
                case 0: case 1: launch_class<2>(ctx, c, P, V, d_tasks, fbl, d_fb_ctr, d_fb_ctr + kExtCtr, nullptr, d_out, sc); break;
                case 2: case 3: launch_class<3>(ctx, c, P, V, d_tasks, fbl, d_fb_ctr, d_fb_ctr + kExtCtr, nullptr, d_out, sc); break;
                case 4: case 5: launch_class<4>(ctx, c, P, V, d_tasks, fbl, d_fb_ctr, d_fb_ctr + kExtCtr, nullptr, d_out, sc); break;
                case 6: case 7: launch_class<5>(ctx, c, P, V, d_tasks, fbl, d_fb_ctr, d_fb_ctr + kExtCtr, nullptr, d_out, sc); break;
                default: launch_class<9>(ctx, c, P, V, d_tasks, fbl, d_fb_ctr, d_fb_ctr + kExtCtr, nullptr, d_out, sc); break;
                }
            }
        } else if (big) {
            int rc = qm_ext2_launch_class(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts ? h_counts[c] : -1, d_out, sc,
                                          scores_fit_bytes);
            if (rc) return rc;
        } else {
            switch (c) {
            case 0: case 1: launch_class<2>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            case 2: case 3: launch_class<3>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            case 4: case 5: launch_class<4>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            case 6: case 7: launch_class<5>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            case 8: launch_class<9>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            default: launch_class<16>(ctx, c, P, V, d_tasks, lst, d_counts, d_cursors, h_counts, d_out, sc); break;
            }
        }
        QM_CUDA(ctx, cudaEventRecord(ctx->ev_join[c], sc));
        QM_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join[c], 0));
    }
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

// Launch a batch of public tasks on `stream`.  Scratch 0: internal tasks + lists + counters.
int qm_extend_launch(qm_ctx *ctx, const qm_opt *opt, const uint8_t *d_seq, const qm_ext_task *d_tasks,
                     int64_t n, qm_ext_result *d_out, cudaStream_t st)
{
    if (n == 0) return QM_OK;
    if (n < 0 || n > 0x7fffffff / kNumClasses) return qm_fail(ctx, QM_ELIMIT, "qm_extend_batch: n_tasks=%lld out of range", (long long)n);
    if (opt->e_ins <= 0 || opt->e_del <= 0) return qm_fail(ctx, QM_EINVAL, "gap extension penalties must be > 0");
    void *p = nullptr;
    const size_t task_bytes = (size_t)n * sizeof(ExtTaskI);
    const size_t list_bytes = (size_t)kNumClasses * n * sizeof(int);
    int rc = qm_scratch_reserve(ctx, 0, task_bytes + 2 * list_bytes + 128 * sizeof(int), &p);
    if (rc) return rc;
    ExtTaskI *itasks = (ExtTaskI *)p;
    int *lists = (int *)((char *)p + task_bytes);
    int *fb_lists = (int *)((char *)p + task_bytes + list_bytes);
    int *ctrs = (int *)((char *)p + task_bytes + 2 * list_bytes);     // [0..9] counts, [16..25] cursors, [32] err, [64..95] fallback
    QM_CUDA(ctx, cudaMemsetAsync(ctrs, 0, 128 * sizeof(int), st));
    const int tpb = 256;
    ext_classify_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>(d_tasks, d_seq, n, itasks, lists, ctrs, d_out, ctrs + 32);
    IndexView V = {};
    return qm_ext_launch_classes(ctx, qm_ext_params(opt), V, itasks, lists, n, ctrs, ctrs + kExtCtr, nullptr, d_out, fb_lists, ctrs + 64, st);
}

extern "C" {

int qm_extend_batch(qm_ctx *ctx, const qm_opt *opt, const uint8_t *d_seq, const qm_ext_task *d_tasks,
                    int64_t n_tasks, qm_ext_result *d_out, void *stream)
{
    if (!ctx || !opt || (n_tasks > 0 && (!d_seq || !d_tasks || !d_out))) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    return qm_extend_launch(ctx, opt, d_seq, d_tasks, n_tasks, d_out, (cudaStream_t)stream);
}

int qm_extend_batch_host(qm_ctx *ctx, const qm_opt *opt, const uint8_t *h_seq, size_t seq_bytes,
                         const qm_ext_task *h_tasks, int64_t n_tasks, qm_ext_result *h_out)
{
    if (!ctx || !opt || n_tasks < 0 || (n_tasks > 0 && (!h_seq || !h_tasks || !h_out))) return QM_EINVAL;
    if (n_tasks == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    void *p = nullptr;
    const size_t seq_al = (seq_bytes + 255) & ~(size_t)255;
    const size_t task_bytes = (size_t)n_tasks * sizeof(qm_ext_task), res_bytes = (size_t)n_tasks * sizeof(qm_ext_result);
    int rc = qm_scratch_reserve(ctx, 1, seq_al + task_bytes + res_bytes, &p);
    if (rc) return rc;
    uint8_t *d_seq = (uint8_t *)p;
    qm_ext_task *d_tasks = (qm_ext_task *)((char *)p + seq_al);
    qm_ext_result *d_out = (qm_ext_result *)((char *)p + seq_al + task_bytes);
    cudaStream_t st = ctx->own_stream;
    QM_CUDA(ctx, cudaMemcpyAsync(d_seq, h_seq, seq_bytes, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(d_tasks, h_tasks, task_bytes, cudaMemcpyHostToDevice, st));
    rc = qm_extend_launch(ctx, opt, d_seq, d_tasks, n_tasks, d_out, st);
    if (rc) return rc;
    QM_CUDA(ctx, cudaMemcpyAsync(h_out, d_out, res_bytes, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n_tasks; ++i)
        if (h_out[i].score == INT32_MIN)
            return qm_fail(ctx, QM_ELIMIT, "task %lld rejected (qlen in [0,%d], tlen>=0, h0>0 required)", (long long)i, QM_EXT_MAX_QLEN);
    return QM_OK;
}

}  // extern "C"
