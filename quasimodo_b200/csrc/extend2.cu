// extend2.cu -- batched ksw_extend2, formulation "T": ONE THREAD PER TASK.
// (bwa 0.7.17 ksw.c:ksw_extend2 as called by bwamem.c:mem_chain2aln; reference call site rules/bwa.smk:15;
// semantics SURVEY.md A.3.)
//
// Each thread runs the reference's row/column loop verbatim on its own task, with the reference's in-place eh[]
// array held in shared memory as two 16-bit planes (h, e; values < 2^16) next to a per-column 16-bit PRMT selector
// that turns the query base into its substitution score in one instruction.  Plane layout: 32-bit word
// (j/2) * T + t holds columns j (low half) and j+1 (high half) of thread t, so lane t always hits bank t%32 --
// conflict free whatever columns the lanes of a warp are at (a plain [column][thread] halfword layout makes
// neighbouring lanes share a bank: 35 % extra wavefronts in the first profile).  16-bit loads/stores keep the
// pack/unpack work off the ALU pipe.  Band trimming, stale cells, z-drop, the to-end score and the band
// retry are the reference's own statements, so exactness needs no argument beyond "same loop".
//
// Why this shape on B200: the extension is bound by the integer ALU pipe (16 lanes/clk/SMSP, DPX included).  The
// warp-per-task kernel (extend.cu) spends ~70 warp-instructions of scan/reduction/control per row; here a cell
// costs ~10 ALU-pipe + ~5 FMA-pipe (IMAD) + 3 LSU instructions and the per-row control is amortised over the 32
// tasks of a warp.  Cell updates use DPX: VIMNMX3 (H = max(M,e,f)), VIADDMNMX.RELU (E and F updates), VIMNMX.
#include "pipeline.cuh"

namespace {

constexpr int kT = 128;                 // threads per block

struct Lut { unsigned lo, hi; };        // 8 score bytes: entries 0..3 = query A,C,G,T, entry 4 = query N

// substitution scores of target base tb against query codes 0..4 as signed bytes (N involved: -1)
__device__ __forceinline__ Lut make_lut(const ExtParams &P, int tb)
{
    Lut L;
    if (tb > 3) { L.lo = 0xffffffffu; L.hi = 0xffffffffu; return L; }
    const unsigned mis = (unsigned)(-P.b) & 0xffu, mat = (unsigned)P.a & 0xffu;
    unsigned v = mis * 0x01010101u;
    v = (v & ~(0xffu << (8 * tb))) | (mat << (8 * tb));
    L.lo = v; L.hi = 0xffffffffu;
    return L;
}

// PTX prmt in its default mode: selector nibble bit 3 replicates the sign of the selected byte, which turns
// "byte q of the LUT" into a sign-extended int in ONE instruction (__byte_perm masks that bit away).
__device__ __forceinline__ int lut_score(const Lut &L, unsigned sel)
{
    int r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(L.lo), "r"(L.hi), "r"(sel));
    return r;
}

struct TState {                         // one in-flight task (registers)
    int tid_out;                        // index of the task (result slot), -1 = idle
    int qlen, tlen, h0, w0, w, end_bonus, tries_left, prev, cells;
    int i, beg, end;
    int mx, mx_i, mx_j, mx_ie, gscore, max_off;
    const uint8_t *q, *t;
    int64_t t0;
    int qstep, tstep;
    int tb_next;                        // target base of the next row, fetched one row ahead
    bool indirect;
};

// BYTES: every score of the batch fits a byte (h0 + qlen*a <= 255, known on the host from the round's counters): eh[j] is
// then ONE 16-bit word h | e << 8 -- 4 instead of 6 bytes of shared memory per column (1.5x the resident warps) and 3
// instead of 5 shared accesses per cell, for 3 more ALU instructions (unpack, pack).
template <int CAP, bool SYM, bool BYTES>
__global__ void __launch_bounds__(kT)
ext2_kernel(ExtParams P, IndexView V, const ExtTaskI *__restrict__ tasks, const int *__restrict__ list,
            const int *__restrict__ count, int *__restrict__ cursor, qm_ext_result *__restrict__ out)
{
    extern __shared__ unsigned short smem_u16[];
    constexpr int PL = (CAP / 2 + 1) * kT * 2;             // halfwords per plane
    unsigned short *HP = smem_u16 + 2 * threadIdx.x;       // HP[IX(j)] = eh[j].h, HP[PL + IX(j)] = eh[j].e,
                                                           // HP[2 PL + IX(j)] = PRMT selector of q[j]
                                                           // (BYTES: HP[IX(j)] = h | e << 8, HP[PL + IX(j)] = selector)
    constexpr int SEL = BYTES ? PL : 2 * PL;
#define EH_SET(j, h) { HP[IX(j)] = (unsigned short)(h); if (!BYTES) HP[PL + IX(j)] = 0; }       /* eh[j] = {h, 0} */
#define EH_ZERO(j) (BYTES ? HP[IX(j)] == 0 : (HP[IX(j)] | HP[PL + IX(j)]) == 0)
#define IX(j) ((((j) >> 1) * (2 * kT)) + ((j) & 1))
    const int n = *count;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    TState S;
    S.tid_out = -1;

    // The lanes of a warp walk the same loop body (refill / row / end-of-try) and reconverge once per row.  A lane
    // whose task is finished does not fetch the next one at once: initialising a task is a serial loop over its
    // query, and 32 lanes refilling one by one would stall the warp's row loop 32 times per task generation.
    // Instead idle lanes wait until kRefill of them are idle (or nobody works) and then refill together.
    constexpr int kRefill = 8;
    bool exhausted = false;
    for (;;) {
        const bool idle = S.tid_out < 0;
        const unsigned idle_m = __ballot_sync(0xffffffffu, idle);
        const unsigned want_m = __ballot_sync(0xffffffffu, idle && !exhausted);
        if (idle_m == 0xffffffffu && want_m == 0u) break;                 // list drained and every lane finished
        // ---- fetch + initialise a task ----
        if (idle && !exhausted && (__popc(want_m) >= kRefill || idle_m == 0xffffffffu)) {
            const int idx = atomicAdd(cursor, 1);
            if (idx >= n) exhausted = true;
            else {
                const int tk = list[idx];
                const ExtTaskI t = tasks[tk];
                S.tid_out = tk;
                S.qlen = t.qlen; S.tlen = t.tlen; S.h0 = t.h0; S.w0 = t.w; S.end_bonus = t.end_bonus;
                S.tries_left = (t.flags & QM_EXT_BAND_RETRY) ? 2 : 1;
                S.prev = (t.flags & QM_EXT_PREV_H0) ? t.h0 : -1;
                S.cells = 0;
                S.q = t.q; S.t = t.t; S.t0 = t.t0; S.qstep = t.qstep; S.tstep = t.tstep;
                S.indirect = (t.flags & QM_EXTI_INDIRECT) != 0;
                S.i = -1;                               // "needs row -1 initialisation"
                S.w = S.w0;
                for (int j = 0; j < S.qlen; ++j) {      // PRMT selectors: byte q of the row's score LUT, sign-extended
                    int c = S.q[(int64_t)j * S.qstep];
                    c = c > 4 ? 4 : c;
                    HP[SEL + IX(j)] = (unsigned short)(c * 0x1111 + 0x8880);
                }
            }
        }
        if (S.tid_out < 0) continue;                    // waiting for a refill batch, or finished
        if (S.i < 0) {
            // row -1 of eh[] (SURVEY.md A.3 first lines) and the band clamp of this try
            const int qlen = S.qlen, h0 = S.h0;
            EH_SET(0, h0)
            int v = h0 > oe_ins ? h0 - oe_ins : 0;
            if (qlen >= 1) EH_SET(1, v)
            int j = 2;
            for (; j <= qlen && v > P.e_ins; ++j) { v -= P.e_ins; EH_SET(j, v) }
            for (; j <= qlen; ++j) EH_SET(j, 0)
            int best = P.a > -1 ? P.a : -1;
            if (-P.b > best) best = -P.b;
            int w = S.w;
            int lim = (int)((double)(qlen * best + S.end_bonus - P.o_ins) / P.e_ins + 1.);
            lim = lim > 1 ? lim : 1;
            w = w < lim ? w : lim;
            lim = (int)((double)(qlen * best + S.end_bonus - P.o_del) / P.e_del + 1.);
            lim = lim > 1 ? lim : 1;
            w = w < lim ? w : lim;
            S.w = w;                                   // clamped band of this try (w_used reports the unclamped one)
            S.mx = h0; S.mx_i = -1; S.mx_j = -1; S.mx_ie = -1; S.gscore = -1; S.max_off = 0;
            S.beg = 0; S.end = qlen; S.i = 0;
            S.tb_next = S.tlen > 0 ? (S.indirect ? qm_ref_base(V, S.t0) : S.t[0]) : 0;
        }
        // ---- one row ----
        bool done = true;
        if (S.i < S.tlen) {
            done = false;
            const int i = S.i, qlen = S.qlen, w = S.w;
            int beg = S.beg, end = S.end;
            const int tb = S.tb_next;
            if (i + 1 < S.tlen) S.tb_next = S.indirect ? qm_ref_base(V, S.t0 + (int64_t)(i + 1) * S.tstep) : S.t[i + 1];
            const Lut L = make_lut(P, tb);
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
            int h1 = 0;
            if (beg == 0) { h1 = S.h0 - (P.o_del + P.e_del * (i + 1)); h1 = h1 > 0 ? h1 : 0; }
            int f = 0;
            int bkey = -1;                 // max over the row's cells of H << 18 | (j - beg): ties -> the later column
            int jr = 0;                    // j - beg
            const int nc = end - beg;
            // one cell of the reference's inner loop at halfword offset O from p; K receives H << 18 | U
#define QM_CELL(O, U, K)                                                                         \
            {                                                                                    \
                int hh, e;                                                                       \
                if (BYTES) { const int he = p[O]; hh = he & 0xff; e = he >> 8; }                 \
                else { hh = p[O]; e = p[PL + (O)]; }                                             \
                const int s = lut_score(L, p[SEL + (O)]);                                        \
                /* h ? h + s : <= 0 (a dead diagonal stays dead); s <= h whenever h > 0 needs a <= 1 */ \
                const int M = hh + min(s, SYM ? hh : hh << 8);                                   \
                const int H = __vimax3_s32(M, e, f);                                             \
                const int td = M - oe_del;                                                       \
                const int en = __viaddmax_s32_relu(e, -P.e_del, td);                             \
                f = __viaddmax_s32_relu(f, -P.e_ins, SYM ? td : M - oe_ins);                     \
                if (BYTES) p[O] = (unsigned short)(en * 256 + h1);                               \
                else { p[O] = (unsigned short)h1; p[PL + (O)] = (unsigned short)en; }            \
                h1 = H;                                                                          \
                K = H * 262144 + (U);                                                            \
            }
            if ((beg & 1) && nc > 0) {     // peel to an even column so that the unrolled offsets are constants
                unsigned short *p = HP + IX(beg);
                int k0;
                QM_CELL(0, 0, k0)
                bkey = k0;
                jr = 1;
            }
            {
                unsigned short *p = HP + IX(beg + jr);
                for (; jr + 3 < nc; jr += 4) {
                    int k0, k1, k2, k3;
                    QM_CELL(0, 0, k0) QM_CELL(1, 1, k1) QM_CELL(2 * kT, 2, k2) QM_CELL(2 * kT + 1, 3, k3)
                    const int m4 = max(__vimax3_s32(k0, k1, k2), k3);
                    bkey = __viaddmax_s32(m4, jr, bkey);
                    p += 4 * kT;
                }
            }
            for (; jr < nc; ++jr) {
                unsigned short *p = HP + IX(beg + jr);
                int k0;
                QM_CELL(0, 0, k0)
                bkey = __viaddmax_s32(k0, jr, bkey);
            }
#undef QM_CELL
#undef IX_UNUSED
            // eh[end] = {h1, 0}; the reference's j is max(beg, end) here
            const int jstop = end > beg ? end : beg;
            EH_SET(jstop, h1)
            if (end > beg) S.cells += end - beg;
            int m = 0, mj = -1;
            if (bkey >= 0) { m = bkey >> 18; mj = beg + (bkey & 0x3ffff); }
            if (jstop == qlen) {
                S.mx_ie = S.gscore > h1 ? S.mx_ie : i;
                S.gscore = S.gscore > h1 ? S.gscore : h1;
            }
            bool stop = (m == 0);
            if (!stop) {
                if (m > S.mx) {
                    S.mx = m; S.mx_i = i; S.mx_j = mj;
                    const int d = abs(mj - i);
                    S.max_off = S.max_off > d ? S.max_off : d;
                } else if (P.zdrop > 0) {
                    const int dr = i - S.mx_i, dc = mj - S.mx_j;
                    if (dr > dc) { if (S.mx - m - (dr - dc) * P.e_del > P.zdrop) stop = true; }
                    else         { if (S.mx - m - (dc - dr) * P.e_ins > P.zdrop) stop = true; }
                }
            }
            if (stop) done = true;
            else {
                // trim to the non-zero span of eh[beg..end]
                int a = beg;
                while (a < end && EH_ZERO(a)) ++a;
                int b = end;
                while (b >= a && EH_ZERO(b)) --b;
                S.beg = a;
                S.end = b + 2 < qlen ? b + 2 : qlen;
                S.i = i + 1;
                if (S.i >= S.tlen) done = true;
            }
        }
        // ---- end of a try: band retry or result ----
        if (done) {
            const int wu = S.w0;                        // unclamped band of this try (what w_used reports)
            // first of two tries: mem_chain2aln's stop rule (score unchanged, or the path stayed inside 3/4 band)
            if (S.tries_left == 2 && !(S.mx == S.prev || S.max_off < (wu >> 1) + (wu >> 2))) {
                S.prev = S.mx; S.tries_left = 1; S.w0 = wu << 1; S.w = S.w0; S.i = -1;
            } else {
                qm_ext_result o;
                o.score = S.mx; o.qle = S.mx_j + 1; o.tle = S.mx_i + 1; o.gtle = S.mx_ie + 1; o.gscore = S.gscore;
                o.max_off = S.max_off; o.w_used = wu; o.cells = S.cells;
                out[S.tid_out] = o;
                S.tid_out = -1;
            }
        }
    }
}

template <int CAP, bool BYTES>
void launch2(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks, const int *d_list,
             const int *d_counts, int *d_cursors, int h_count, qm_ext_result *d_out, cudaStream_t st)
{
    const size_t smem = (size_t)(BYTES ? 2 : 3) * (CAP / 2 + 1) * kT * 2 * 2;
    static bool attr_set[64] = {};              // per device: function attributes belong to a device's context, and one
                                                // process may drive several GPUs (qm_driver --gpus)
    if (!attr_set[ctx->device & 63]) {
        cudaFuncSetAttribute(ext2_kernel<CAP, true, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(ext2_kernel<CAP, false, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[ctx->device & 63] = true;
    }
    int per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (h_count >= 0) {
        const int64_t need = (h_count + kT - 1) / kT;
        if (need < blocks) blocks = need;
    }
    if (blocks < 1) return;
    // SYM = the standard scheme: equal gap costs (one subtraction serves E and F) and match score 1 (see QM_CELL)
    const bool sym = P.o_del == P.o_ins && P.e_del == P.e_ins && P.a == 1;
    if (sym) ext2_kernel<CAP, true, BYTES><<<(unsigned)blocks, kT, smem, st>>>(P, V, d_tasks, d_list, d_counts + cls, d_cursors + cls, d_out);
    else ext2_kernel<CAP, false, BYTES><<<(unsigned)blocks, kT, smem, st>>>(P, V, d_tasks, d_list, d_counts + cls, d_cursors + cls, d_out);
}

}  // namespace

int qm_ext2_launch_class(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks,
                         const int *d_list, const int *d_counts, int *d_cursors, int h_count,
                         qm_ext_result *d_out, cudaStream_t st, bool bytes)
{
#define QM_L2(CAPV) { if (bytes) launch2<CAPV, true>(ctx, cls, P, V, d_tasks, d_list, d_counts, d_cursors, h_count, d_out, st); \
                      else launch2<CAPV, false>(ctx, cls, P, V, d_tasks, d_list, d_counts, d_cursors, h_count, d_out, st); }
    switch (cls) {
    case 0: QM_L2(16) break;
    case 1: QM_L2(32) break;
    case 2: QM_L2(48) break;
    case 3: QM_L2(64) break;
    case 4: QM_L2(80) break;
    case 5: QM_L2(96) break;
    case 6: QM_L2(112) break;
    case 7: QM_L2(128) break;
    case 8: QM_L2(256) break;
    default: return qm_fail(ctx, QM_EINVAL, "qm_ext2_launch_class: class %d has no thread-per-task kernel", cls);
    }
#undef QM_L2
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}
