// extend3.cu -- batched ksw_extend2, formulation "P2": the device side of ext3_core.cuh (ONE THREAD PER PAIR OF TASKS,
// both tasks in the halves of s16x2 DPX words; bwa 0.7.17 ksw.c:ksw_extend2 as called by bwamem.c:mem_chain2aln,
// reference call site rules/bwa.smk:15, semantics SURVEY.md A.3).  The per-thread statements live in ext3_core.cuh, which
// the CPU test suite compiles for the host and checks against the oracle; this file adds what only exists on the device:
// the shared-memory planes, the target fetch, the persistent grid with its warp-batched refill, the launch shapes.
//
// Why this shape on B200: the extension is bound by the integer ALU pipe (16 lanes/clk/SMSP, DPX included).  The scalar
// kernel (extend2.cu) spends ~11 ALU-pipe instructions per cell plus ~8 of per-row control; here a PAIR of cells costs ~9.5
// and the per-row control of two tasks is one instruction stream.  Shared memory per thread: 12 bytes per column (64-bit
// {h2, e2} + 32-bit query codes) for two tasks; block sizes are chosen per class so that a few blocks fit an SM.
#include <stdlib.h>
#include <type_traits>
#include "pipeline.cuh"
#include "ext3_core.cuh"

namespace {

// The thread's columns in shared memory, word j of a plane at [j][thread] (a warp's accesses hit consecutive banks whatever
// column each lane is at).  Two layouts:
//   SmemWide   16-bit cells: a 64-bit word {h2, e2} and a 32-bit query word per column -- 12 bytes per column and task pair,
//              any score up to the kernel's limit, no unpacking;
//   SmemNarrow byte cells: one 32-bit word (h | e << 8 in each task's half) and a 16-bit query word -- 6 bytes per column and
//              pair, i.e. TWICE the resident warps, for three more ALU instructions per cell pair (two unpacks, one query
//              expansion; the re-pack is an IMAD on the other pipe).  Needs scores <= 255 (the kernel's limit anyway) and
//              a + b <= 16.  With one to two warps per scheduler the wide layout leaves the ALU pipe idle most cycles
//              (profiles/r02_summary.md), so the narrow one is the default where the scores allow it.
template <int KT>
struct SmemWide {
    static constexpr bool kNarrow = false;
    struct Raw { uint2 he; unsigned q; };
    uint2 *ehp;
    uint32_t *qp;
    static constexpr size_t bytes_per_thread(int cap) { return (size_t)(cap + 1) * 8 + (size_t)(cap + kE3Pad) * 4; }
    __device__ __forceinline__ void bind(void *smem, int cap, int t) { ehp = (uint2 *)smem + t; qp = (uint32_t *)((uint2 *)smem + (cap + 1) * KT) + t; }
    __device__ __forceinline__ void clear(int cap) { for (int j = 0; j <= cap; ++j) ehp[j * KT] = make_uint2(0u, 0u); for (int j = 0; j < cap + kE3Pad; ++j) qp[j * KT] = 0u; }
    __device__ __forceinline__ Raw raw(int j) const { Raw r; r.he = ehp[j * KT]; r.q = qp[j * KT]; return r; }
    static __device__ __forceinline__ void unpack(const Raw &r, unsigned &h2, unsigned &e2, unsigned &q2) { h2 = r.he.x; e2 = r.he.y; q2 = r.q; }
    __device__ __forceinline__ void put(int j, unsigned h2, unsigned e2) { ehp[j * KT] = make_uint2(h2, e2); }
    __device__ __forceinline__ void set_he(int j, int X, int h, int e) { ((uint16_t *)&ehp[j * KT].x)[X] = (uint16_t)h; ((uint16_t *)&ehp[j * KT].y)[X] = (uint16_t)e; }
    __device__ __forceinline__ bool zero(int j, int X) const { return (((const uint16_t *)&ehp[j * KT].x)[X] | ((const uint16_t *)&ehp[j * KT].y)[X]) == 0; }
    __device__ __forceinline__ void set_q(int j, int X, unsigned code) { ((uint16_t *)&qp[j * KT])[X] = (uint16_t)code; }
};

template <int KT>
struct SmemNarrow {
    static constexpr bool kNarrow = true;
    struct Raw { unsigned he; unsigned q; };
    uint32_t *ehp;
    uint16_t *qp;
    static constexpr size_t bytes_per_thread(int cap) { return (size_t)(cap + 1) * 4 + (size_t)(cap + kE3Pad) * 2; }
    __device__ __forceinline__ void bind(void *smem, int cap, int t) { ehp = (uint32_t *)smem + t; qp = (uint16_t *)((uint32_t *)smem + (cap + 1) * KT) + t; }
    __device__ __forceinline__ void clear(int cap) { for (int j = 0; j <= cap; ++j) ehp[j * KT] = 0u; for (int j = 0; j < cap + kE3Pad; ++j) qp[j * KT] = 0; }
    __device__ __forceinline__ Raw raw(int j) const { Raw r; r.he = ehp[j * KT]; r.q = qp[j * KT]; return r; }
    static __device__ __forceinline__ void unpack(const Raw &r, unsigned &h2, unsigned &e2, unsigned &q2)
    {
        h2 = r.he & 0x00ff00ffu; e2 = __byte_perm(r.he, 0u, 0x4341); q2 = __byte_perm(r.q, 0u, 0x4140);
    }
    // (every stored value is <= 255: an active task's scores are, a task-less half only ever sees what was stored)
    __device__ __forceinline__ void put(int j, unsigned h2, unsigned e2) { ehp[j * KT] = e2 * 256u + h2; }
    __device__ __forceinline__ void set_he(int j, int X, int h, int e) { ((uint16_t *)&ehp[j * KT])[X] = (uint16_t)((h & 0xff) | (e & 0xff) << 8); }
    __device__ __forceinline__ bool zero(int j, int X) const { return ((const uint16_t *)&ehp[j * KT])[X] == 0; }
    __device__ __forceinline__ void set_q(int j, int X, unsigned code) { ((uint8_t *)&qp[j * KT])[X] = (uint8_t)code; }
};

struct DevTgt {                             // where the two tasks' target rows come from: base of row i = p[i * step] ^ flip
    const uint8_t *p[2];
    int step[2];
    int flip[2];                            // 3 on the reverse-complement strand of the doubled coordinates, else 0
    __device__ __forceinline__ int raw(int X, int i) const { return p[X][(int64_t)i * step[X]]; }
    __device__ __forceinline__ int decode(int X, int c) const { c ^= flip[X]; return c > 4 ? 4 : c; }
    // a task's window never crosses the strand boundary (mem_chain2aln clamps rmax[] to one strand), so the strand test of
    // qm_ref_base is made once per task instead of once per row
    __device__ __forceinline__ void set(int X, const IndexView &V, const ExtTaskI &t)
    {
        if (!(t.flags & QM_EXTI_INDIRECT)) { p[X] = t.t; step[X] = 1; flip[X] = 0; }
        else if (t.t0 < V.l_pac) { p[X] = V.refb + t.t0; step[X] = t.tstep; flip[X] = 0; }
        else { p[X] = V.refb + (2 * V.l_pac - 1 - t.t0); step[X] = -t.tstep; flip[X] = 3; }
    }
};

struct DevQry {
    const uint8_t *q;
    int qstep;
    __device__ __forceinline__ int code(int j) const { const int c = q[(int64_t)j * qstep]; return c > 4 ? 4 : c; }
};

template <int CAP, int KT, bool SYM, bool NARROW>
__global__ void __launch_bounds__(KT)
ext3_kernel(ExtParams P, IndexView V, const ExtTaskI *__restrict__ tasks, const int *__restrict__ list,
            const int *__restrict__ count, int *__restrict__ cursor, qm_ext_result *__restrict__ out,
            int *__restrict__ fb_list, int *__restrict__ fb_count)
{
    extern __shared__ uint2 smem_u2[];
    typename std::conditional<NARROW, SmemNarrow<KT>, SmemWide<KT>>::type mem;
    mem.bind(smem_u2, CAP, threadIdx.x);
    // dead storage must hold small non-negative values (a half without a task rides along on its partner's columns)
    mem.clear(CAP);
    const E3Scores S = {P.a, P.b, P.o_del, P.e_del, P.o_ins, P.e_ins, P.zdrop};
    const E3Consts K = e3_consts(S);
    const int n = *count;
    E3Half A, B;
    A.tk = B.tk = -1; A.phase = B.phase = 0;
    DevTgt tgt;
    tgt.p[0] = tgt.p[1] = nullptr; tgt.step[0] = tgt.step[1] = 0; tgt.flip[0] = tgt.flip[1] = 0;

    auto load = [&](E3Half &H, int X, int tk) {
        const ExtTaskI t = tasks[tk];
        if (!e3_task_ok(S, t.qlen, t.h0, CAP)) { fb_list[atomicAdd(fb_count, 1)] = tk; return; }      // the scalar kernel takes it
        H.tk = tk; H.phase = 1;
        H.qlen = t.qlen; H.tlen = t.tlen; H.h0 = t.h0; H.w0 = t.w; H.w = t.w; H.end_bonus = t.end_bonus;
        H.tries_left = (t.flags & QM_EXT_BAND_RETRY) ? 2 : 1;
        H.prev = (t.flags & QM_EXT_PREV_H0) ? t.h0 : -1;
        H.cells = 0;
        tgt.set(X, V, t);
        DevQry qry = {t.q, t.qstep};
        e3_load_query(K, t.qlen, X, mem, qry);
    };
    auto finish = [&](E3Half &H) {
        E3Result r;
        if (e3_end_try(H, &r)) {
            qm_ext_result o;
            o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore;
            o.max_off = r.max_off; o.w_used = r.w_used; o.cells = r.cells;
            out[H.tk] = o;
            H.tk = -1;
        }
    };

    // The lanes of a warp walk the same loop body (refill / start of a try / row) and reconverge once per row.  A lane whose
    // two tasks are finished does not fetch the next pair at once: loading a task is a serial loop over its query, and 32
    // lanes refilling one by one would stall the warp's row loop 32 times per generation.  Idle lanes wait until kRefill of
    // them are idle (or nobody works) and refill together.
    constexpr int kRefill = 8;
    bool exhausted = false;
    for (;;) {
        const bool idle = A.phase == 0 && B.phase == 0;
        const unsigned idle_m = __ballot_sync(0xffffffffu, idle);
        const unsigned want_m = __ballot_sync(0xffffffffu, idle && !exhausted);
        if (idle_m == 0xffffffffu && want_m == 0u) break;
        if (idle && !exhausted && (__popc(want_m) >= kRefill || idle_m == 0xffffffffu)) {
            const int idx = atomicAdd(cursor, 2);
            if (idx >= n) exhausted = true;
            else {
                load(A, 0, list[idx]);
                if (idx + 1 < n) load(B, 1, list[idx + 1]);
            }
        }
        if (A.phase == 1) e3_start_try(K, A, 0, mem, tgt);
        if (B.phase == 1) e3_start_try(K, B, 1, mem, tgt);
        if (A.phase == 0 && B.phase == 0) continue;
        bool dA, dB;
        e3_row<SYM>(K, A, B, mem, tgt, dA, dB);
        if (dA) finish(A);
        if (dB) finish(B);
    }
}

template <int CAP, int KT, bool NARROW>
void launch3(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks, const int *list,
             const int *d_counts, int *d_cursors, int h_count, qm_ext_result *d_out, int *fb,
             int *d_fb_ctr, cudaStream_t st)
{
    const size_t smem = (size_t)KT * (NARROW ? SmemNarrow<KT>::bytes_per_thread(CAP) : SmemWide<KT>::bytes_per_thread(CAP));
    static bool attr_set[64] = {};              // per device: function attributes belong to a device's context, and one
                                                // process may drive several GPUs (qm_driver --gpus)
    if (!attr_set[ctx->device & 63]) {
        cudaFuncSetAttribute(ext3_kernel<CAP, KT, true, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(ext3_kernel<CAP, KT, false, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[ctx->device & 63] = true;
    }
    int per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 32) per_sm = 32;
    if (per_sm * KT > 2048) per_sm = 2048 / KT;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (h_count >= 0) {
        const int64_t need = ((h_count + 1) / 2 + KT - 1) / KT;
        if (need < blocks) blocks = need;
    }
    if (blocks < 1) return;
    const bool sym = P.o_del == P.o_ins && P.e_del == P.e_ins && P.a == 1;
    if (sym) ext3_kernel<CAP, KT, true, NARROW><<<(unsigned)blocks, KT, smem, st>>>(P, V, d_tasks, list, d_counts + cls, d_cursors + cls, d_out, fb, d_fb_ctr + cls);
    else ext3_kernel<CAP, KT, false, NARROW><<<(unsigned)blocks, KT, smem, st>>>(P, V, d_tasks, list, d_counts + cls, d_cursors + cls, d_out, fb, d_fb_ctr + cls);
}

}  // namespace

bool qm_ext3_scores_ok(const ExtParams &P)
{
    const E3Scores S = {P.a, P.b, P.o_del, P.e_del, P.o_ins, P.e_ins, P.zdrop};
    return e3_scores_ok(S);
}

// one class (0..8) on the packed kernel (d_list = the class's task indices); tasks it cannot hold (scores above 255) are
// appended to the class's fallback list d_fb_list (count d_fb_ctr[cls]) for the scalar kernel.  h_count < 0: unknown on the host.
int qm_ext3_launch_class(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks,
                         const int *d_list, const int *d_counts, int *d_cursors, int h_count,
                         qm_ext_result *d_out, int *d_fb_list, int *d_fb_ctr, cudaStream_t st)
{
    // byte planes wherever the scoring scheme fits them (QM_EXT3_NARROW=0: 16-bit planes everywhere, for A/B measurements)
    static const bool want_narrow = !(getenv("QM_EXT3_NARROW") && atoi(getenv("QM_EXT3_NARROW")) == 0);
    const E3Scores S = {P.a, P.b, P.o_del, P.e_del, P.o_ins, P.e_ins, P.zdrop};
    const bool narrow = want_narrow && e3_scores_ok_narrow(S);
#define QM_L3(CAPV, KTW, KTN) { if (narrow) launch3<CAPV, KTN, true>(ctx, cls, P, V, d_tasks, d_list, d_counts, d_cursors, h_count, d_out, d_fb_list, d_fb_ctr, st); \
                                else launch3<CAPV, KTW, false>(ctx, cls, P, V, d_tasks, d_list, d_counts, d_cursors, h_count, d_out, d_fb_list, d_fb_ctr, st); }
    switch (cls) {
    case 0: QM_L3(16, 64, 64) break;
    case 1: QM_L3(32, 64, 64) break;
    case 2: QM_L3(48, 64, 64) break;
    case 3: QM_L3(64, 64, 64) break;
    case 4: QM_L3(80, 32, 64) break;
    case 5: QM_L3(96, 32, 64) break;
    case 6: QM_L3(112, 32, 64) break;
    case 7: QM_L3(128, 32, 64) break;
    case 8: QM_L3(256, 32, 32) break;
    default: return qm_fail(ctx, QM_EINVAL, "qm_ext3_launch_class: class %d has no packed kernel", cls);
    }
#undef QM_L3
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}
