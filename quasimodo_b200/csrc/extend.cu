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
#include "pipeline.cuh"

namespace {

constexpr int kNumClasses = kExtClasses;
constexpr int kThreadPerTaskMin = 16384;      // tasks in a class below which the warp-per-task kernel is used

// where a task's bases come from: explicit byte strings (public C-ABI tasks) or the read batch /
// reference index in place (pipeline tasks; nothing is materialised in HBM)
struct SeqFetch {
    const uint8_t *q, *t;
    int64_t t0;
    int qstep, tstep;
    bool indirect;
    const IndexView *V;
    __device__ __forceinline__ int qbase(int j) const { return q[(int64_t)j * qstep]; }
    __device__ __forceinline__ int tbase(int i) const
    {
        return indirect ? qm_ref_base(*V, t0 + (int64_t)i * tstep) : t[i];
    }
};

struct ExtState {
    int score, qle, tle, gtle, gscore, max_off, cells;
};

// One ksw_extend2 call executed by one warp.  All lanes return the same ExtState.
template <int C>
__device__ __forceinline__ ExtState ext_run(const ExtParams &P, const SeqFetch &S, int qlen, int tlen, int h0, int w,
                                            int end_bonus, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int j0 = lane * C;

    int h[C], e[C], qc[C], mis[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = j0 + k;
        // row -1 of eh[].h (SURVEY.md A.3 first three lines), closed form
        int v = 0;
        if (j == 0) v = h0;
        else if (j <= qlen) {
            const int vj = h0 - oe_ins - (j - 1) * P.e_ins;
            if (j == 1) v = vj > 0 ? vj : 0;
            else v = (vj + P.e_ins > P.e_ins) ? vj : 0;
        }
        h[k] = v;
        e[k] = 0;
        const int c = (j < qlen) ? S.qbase(j) : 4;
        qc[k] = c;
        mis[k] = (c > 3) ? -1 : -P.b;
    }

    {   // band cannot usefully exceed what the scores can pay for (doubles, as the reference)
        int best = P.a > -1 ? P.a : -1;
        if (-P.b > best) best = -P.b;
        int lim = (int)((double)(qlen * best + end_bonus - P.o_ins) / P.e_ins + 1.);
        lim = lim > 1 ? lim : 1;
        w = w < lim ? w : lim;
        lim = (int)((double)(qlen * best + end_bonus - P.o_del) / P.e_del + 1.);
        lim = lim > 1 ? lim : 1;
        w = w < lim ? w : lim;
    }

    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen, cells = 0;

    int tb_next = tlen > 0 ? S.tbase(0) : 0;
    for (int i = 0; i < tlen; ++i) {
        const int tb = tb_next;
        if (i + 1 < tlen) tb_next = S.tbase(i + 1);

        if (beg < i - w) {
            // columns that fall out of the band are never read again by the reference; zero them so
            // that they behave like the reference's trimmed (dead) columns.
            beg = i - w;
#pragma unroll
            for (int k = 0; k < C; ++k)
                if (j0 + k < beg) { h[k] = 0; e[k] = 0; }
        }
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) {
            h1_init = h0 - (P.o_del + P.e_del * (i + 1));
            h1_init = h1_init > 0 ? h1_init : 0;
        }
        const int na = end - j0;       // columns k < na of this lane are inside [.., end)
        const bool tN = tb > 3;

        // ---- pass 1: M, E(i+1,.), lane-local F chain with zero carry-in ----
        int M[C], fin[C], en[C];
        int f = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            int s = (qc[k] == tb) ? P.a : mis[k];
            if (tN) s = -1;
            const int hk = h[k];
            const int m = hk + min(s, hk);          // == hk ? hk + s : <=0  (a dead diagonal stays dead)
            M[k] = m;
            en[k] = __viaddmax_s32_relu(e[k], -P.e_del, m - oe_del);
            fin[k] = f;
            f = __viaddmax_s32_relu(f, -P.e_ins, m - oe_ins);
        }
        // ---- warp max-plus scan of the F carry ----
        int carry = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, carry, d);
            carry = __viaddmax_s32(o, -d * C * P.e_ins, carry);   // lanes < d get their own value back: harmless
        }
        int Fin = __shfl_up_sync(FULL, carry, 1);
        if (lane == 0) Fin = 0;

        // ---- pass 2: H(i,.), row statistics ----
        int H[C];
        int key = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            int hk = __vimax3_s32(M[k], e[k], fin[k]);
            hk = __viaddmax_s32(Fin, -k * P.e_ins, hk);
            hk = (k < na) ? hk : 0;
            H[k] = hk;
            key = max(key, (hk << 9) | (j0 + k));
        }
        // ---- commit eh[]: h shifts one column right, e in place; only indices <= end are written ----
        int hleft = __shfl_up_sync(FULL, H[C - 1], 1);
        if (lane == 0) hleft = h1_init;
        int nzlast = -1, nzfirst = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int hl = (k == 0) ? hleft : H[k - 1];
            if (k <= na) h[k] = hl;
            e[k] = (k < na) ? en[k] : ((k == na) ? 0 : e[k]);
            const bool nz = (k <= na) && ((h[k] | e[k]) != 0);
            nzlast = nz ? (j0 + k) : nzlast;
            nzfirst = (nz && nzfirst == 0x7fffffff) ? (j0 + k) : nzfirst;
        }
        const int kmax = __reduce_max_sync(FULL, key);
        const int m = kmax >> 9, mj = kmax & 511;

        if (end > beg) cells += end - beg;
        const int jstop = end > beg ? end : beg;
        if (jstop == qlen) {
            int h1 = h1_init;
            if (end > beg) {
                int v = 0;
#pragma unroll
                for (int k = 0; k < C; ++k) v = (j0 + k == end - 1) ? H[k] : v;
                h1 = __reduce_max_sync(FULL, v);
            }
            mx_ie = gscore > h1 ? mx_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > mx) {
            mx = m; mx_i = i; mx_j = mj;
            const int d = abs(mj - i);
            max_off = max_off > d ? max_off : d;
        } else if (P.zdrop > 0) {
            const int dr = i - mx_i, dc = mj - mx_j;
            if (dr > dc) { if (mx - m - (dr - dc) * P.e_del > P.zdrop) break; }
            else         { if (mx - m - (dc - dr) * P.e_ins > P.zdrop) break; }
        }
        // ---- trim to the non-zero span of eh[beg..end] ----
        const int jl = __reduce_max_sync(FULL, nzlast);
        const int jf = __reduce_min_sync(FULL, nzfirst);
        beg = jf;
        end = jl + 2 < qlen ? jl + 2 : qlen;
    }
    ExtState r;
    r.score = mx; r.qle = mx_j + 1; r.tle = mx_i + 1; r.gtle = mx_ie + 1; r.gscore = gscore;
    r.max_off = max_off; r.cells = cells;
    return r;
}

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
void launch_class(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks, const int *d_lists,
                  int64_t list_stride, const int *d_counts, int *d_cursors, const int *h_counts, qm_ext_result *d_out,
                  cudaStream_t st)
{
    int64_t warps = (int64_t)ctx->sm_count * 64;          // persistent: up to 16 blocks of 4 warps per SM
    if (h_counts) {
        if (h_counts[cls] == 0) return;
        if (h_counts[cls] < warps) warps = h_counts[cls];
    }
    unsigned blocks = (unsigned)((warps + 3) / 4);
    if (blocks > (unsigned)ctx->sm_count) blocks = ((blocks + ctx->sm_count - 1) / ctx->sm_count) * ctx->sm_count;
    ext_kernel<C><<<blocks, 128, 0, st>>>(P, V, d_tasks, d_lists + cls * list_stride, d_counts + cls, d_cursors + cls, d_out);
}

}  // namespace

int qm_ext_launch_classes(qm_ctx *ctx, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks,
                          const int *d_lists, int64_t list_stride, const int *d_counts, int *d_cursors,
                          const int *h_counts, qm_ext_result *d_out, cudaStream_t st)
{
    // A class with many tasks goes to the thread-per-task kernel (extend2.cu: high throughput, but one task is a
    // long serial chain, ~0.3 ms); a class with few tasks (the tail rounds of mem_chain2aln, where only reads with
    // many chains are still active) goes to the warp-per-task kernel below (low latency).  Class 5 (qlen > 256)
    // always does.  Without host-side counts (public qm_extend_batch) classes 0..4 use the thread-per-task kernel.
    int big[kExtClasses] = {1, 1, 1, 1, 1, 0};
    if (h_counts)
        for (int c = 0; c < 5; ++c) big[c] = h_counts[c] >= kThreadPerTaskMin;
    int hc2[kExtClasses];
    for (int c = 0; c < kExtClasses; ++c) hc2[c] = big[c] ? (h_counts ? h_counts[c] : 1) : 0;
    int rc = qm_ext2_launch_classes(ctx, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, hc2, h_counts != nullptr, d_out, st);
    if (rc) return rc;
    if (!big[0]) launch_class<2>(ctx, 0, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
    if (!big[1]) launch_class<3>(ctx, 1, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
    if (!big[2]) launch_class<4>(ctx, 2, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
    if (!big[3]) launch_class<5>(ctx, 3, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
    if (!big[4]) launch_class<9>(ctx, 4, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
    launch_class<16>(ctx, 5, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_counts, d_out, st);
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
    int rc = qm_scratch_reserve(ctx, 0, task_bytes + list_bytes + 64 * sizeof(int), &p);
    if (rc) return rc;
    ExtTaskI *itasks = (ExtTaskI *)p;
    int *lists = (int *)((char *)p + task_bytes);
    int *ctrs = (int *)((char *)p + task_bytes + list_bytes);     // [0..5] counts, [8..13] cursors, [16] err
    QM_CUDA(ctx, cudaMemsetAsync(ctrs, 0, 64 * sizeof(int), st));
    const int tpb = 256;
    ext_classify_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>(d_tasks, d_seq, n, itasks, lists, ctrs, d_out, ctrs + 16);
    IndexView V = {};
    return qm_ext_launch_classes(ctx, qm_ext_params(opt), V, itasks, lists, n, ctrs, ctrs + 8, nullptr, d_out, st);
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
