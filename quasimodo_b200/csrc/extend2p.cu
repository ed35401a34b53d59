// extend2p.cu -- batched ksw_extend2, formulation "P": ONE THREAD PER PAIR OF TASKS, both tasks in one s16x2 word.
// (bwa 0.7.17 ksw.c:ksw_extend2 as called by bwamem.c:mem_chain2aln; reference call site rules/bwa.smk:15;
// semantics SURVEY.md A.3.)
//
// Two independent tasks A and B share a thread: A lives in the low, B in the high 16 bits of every DP word, and the
// thread walks row i of both at once.  Every extension starts at the origin and follows the main diagonal, so at the
// same row the two tasks' column ranges [beg, end) overlap almost completely; over the overlap one DPX instruction
// updates a cell of each task (VIMNMX3.S16x2, VIADDMNMX.S16x2.RELU, VIADD.16x2, one PRMT that looks up both
// substitution scores).  Columns only one task visits run the same code with a half-word blend on the store, so the
// other task's cells keep their (stale) values exactly as the reference's in-place eh[] array does.
//
// Shared memory per thread: ONE 32-bit word per column holding the reference's eh[j] of both tasks as four bytes
// {hA, eA, hB, eB} ([column][thread]: lane t always hits bank t, conflict free whatever column each lane is at) and
// one 32-bit PRMT-selector word per column pair -- 3 bytes per column and task against 6 in extend2.cu, so twice the
// tasks are resident per SM.  Per cell pair: 1 LDS + 1 STS (+ half a selector LDS), ~10 ALU-pipe instructions
// (extend2.cu: 5 shared accesses and ~9 ALU instructions per SINGLE cell); the byte pack of the store and the row
// maximum keys are IMADs on the FMA pipe.  Bytes need H <= 255; row maxima are tracked as packed unsigned keys
// H << 7 | column (columns <= 127).  Tasks outside that (h0 + qlen*a > 255, qlen > 127) or with an N in the query
// (the 8 LUT bytes of the PRMT hold the A/C/G/T scores of both tasks, no room for N) are handed to the scalar kernel
// through a per-class fallback list.  Band trimming, stale cells, z-drop, the to-end score and the band retry are the
// reference's statements per task, as in extend2.cu.
#include "pipeline.cuh"

namespace {

constexpr int kT = 128;                 // threads per block

__device__ __forceinline__ unsigned add2(unsigned a, unsigned b) { unsigned r; asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned min2(unsigned a, unsigned b) { unsigned r; asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned maxu2(unsigned a, unsigned b) { unsigned r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned s) { unsigned r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
__device__ __forceinline__ unsigned pack2(int v) { return ((unsigned)v & 0xffffu) * 0x10001u; }

// scores of target base tb against query A,C,G,T as signed bytes (a target N scores -1 against everything)
__device__ __forceinline__ unsigned make_lut4(const ExtParams &P, int tb)
{
    if (tb > 3) return 0xffffffffu;
    const unsigned mis = (unsigned)(-P.b) & 0xffu, mat = (unsigned)P.a & 0xffu;
    const unsigned v = mis * 0x01010101u;
    return (v & ~(0xffu << (8 * tb))) | (mat << (8 * tb));
}

struct Half {                           // one task of the pair (registers)
    int tid_out;                        // result slot, -1 = no task in this half
    int run;                            // 1 = rows in progress, 0 = idle / finished / waiting for its second try
    int retry;                          // 1 = first try done, second try (doubled band) waits for the partner
    int qlen, tlen, h0, w0, w, end_bonus, tries_left, prev, cells;
    int beg, end;
    int mx, mx_i, mx_j, mx_ie, gscore, max_off;
    const uint8_t *t;
    int64_t t0;
    int tstep;
    int tb_next;
    bool indirect;
};

template <int CAP, bool SYM>
__global__ void __launch_bounds__(kT)
ext2p_kernel(ExtParams P, IndexView V, const ExtTaskI *__restrict__ tasks, const int *__restrict__ list,
             const int *__restrict__ count, int *__restrict__ cursor, qm_ext_result *__restrict__ out,
             int *__restrict__ fb_list, int *__restrict__ fb_count)
{
    extern __shared__ unsigned smem_u32[];
    constexpr int PLW = (CAP + 1) * kT;                     // words of the eh plane
    unsigned *HW = smem_u32 + threadIdx.x;                  // HW[j * kT] = bytes {hA, eA, hB, eB} of column j
    unsigned *SW = HW + PLW;                                // SW[(j >> 1) * kT] = selectors of columns j (low) and j+1 (high)
    unsigned short *EH16 = (unsigned short *)HW;            // task X of column j: EH16[2 * j * kT + X] = h | e << 8
    unsigned short *S16 = (unsigned short *)SW;             // selector of column j: S16[(j >> 1) * 2 * kT + (j & 1)]
    const int n = *count;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const unsigned noe_del2 = pack2(-oe_del), noe_ins2 = pack2(-oe_ins), ned2 = pack2(-P.e_del), nei2 = pack2(-P.e_ins);
    unsigned kc1 = 0x10001u, kc2 = 0x20002u, kc3 = 0x30003u;    // column offsets inside a group of 4, kept in registers so that
    asm volatile("" : "+r"(kc1), "+r"(kc2), "+r"(kc3));         // the key is one IMAD (FMA pipe), not an ALU-pipe LEA
    Half T[2];
    T[0].tid_out = T[1].tid_out = -1;
    T[0].run = T[1].run = 0; T[0].retry = T[1].retry = 0;
    int i = 0;                                              // current row of both halves

    constexpr int kRefill = 8;
    bool exhausted = false;
    int pend = -1;                                          // a fetched task with an N, waiting for this thread's low half
    for (;;) {
        const bool idle = !T[0].run && !T[1].run && !T[0].retry && !T[1].retry;
        const unsigned idle_m = __ballot_sync(0xffffffffu, idle);
        const unsigned want_m = __ballot_sync(0xffffffffu, idle && (!exhausted || pend >= 0));
        if (idle_m == 0xffffffffu && want_m == 0u) break;
        // ---- fetch two tasks and build the selector plane.  A query with an N cannot share the PRMT's 8 LUT bytes with
        // a partner (A/C/G/T of both tasks fill them): such a task runs ALONE in the low half, its N columns select
        // the partner's LUT, which is all -1 while the high half is empty.  A second task that turns out to have an N
        // is kept for this thread's next refill.  Tasks the byte arithmetic cannot hold go to the fallback list. ----
        if (idle && (!exhausted || pend >= 0) && (__popc(want_m) >= kRefill || idle_m == 0xffffffffu)) {
            const uint8_t *q = nullptr;
            int qstep = 0;
            auto load_half = [&](int X, int tk) -> bool {
                Half &S = T[X];
                const ExtTaskI t = tasks[tk];
                if (!(t.h0 + t.qlen * P.a <= 255 && t.qlen <= 127 && t.qlen <= CAP)) { fb_list[atomicAdd(fb_count, 1)] = tk; return false; }
                S.tid_out = tk;
                S.qlen = t.qlen; S.tlen = t.tlen; S.h0 = t.h0; S.w0 = t.w; S.end_bonus = t.end_bonus;
                S.tries_left = (t.flags & QM_EXT_BAND_RETRY) ? 2 : 1;
                S.prev = (t.flags & QM_EXT_PREV_H0) ? t.h0 : -1;
                S.cells = 0;
                S.t = t.t; S.t0 = t.t0; S.tstep = t.tstep;
                S.indirect = (t.flags & QM_EXTI_INDIRECT) != 0;
                S.w = S.w0;
                S.retry = 1;                                // "start a try at the next opportunity"
                q = t.q; qstep = t.qstep;
                return true;
            };
            T[0].tid_out = T[1].tid_out = -1;
            int tkA = pend;
            pend = -1;
            if (tkA < 0) {
                const int idx = atomicAdd(cursor, 1);
                if (idx < n) tkA = list[idx]; else exhausted = true;
            }
            bool n_in_a = false;
            if (tkA >= 0 && load_half(0, tkA)) {
                for (int j = 0; j < T[0].qlen; ++j) {
                    const unsigned c = q[(int64_t)j * qstep];
                    if (c > 3u) n_in_a = true;
                    const unsigned lo = c > 3u ? 0xc4u : (c | (c | 8u) << 4);
                    S16[(j >> 1) * (2 * kT) + (j & 1)] = (unsigned short)(lo | 0xc400u);
                }
            }
            if (tkA >= 0 && !n_in_a && !exhausted) {
                const int idx = atomicAdd(cursor, 1);
                if (idx >= n) exhausted = true;
                else {
                    const int tkB = list[idx];
                    if (load_half(1, tkB)) {
                        bool n_in_b = false;
                        for (int j = 0; j < T[1].qlen; ++j) {
                            const unsigned c = q[(int64_t)j * qstep];
                            if (c > 3u) { n_in_b = true; break; }
                            const unsigned hi = (c + 4u) | (c + 12u) << 4;
                            unsigned short *sp = &S16[(j >> 1) * (2 * kT) + (j & 1)];
                            *sp = (unsigned short)((*sp & 0xffu) | hi << 8);
                        }
                        if (n_in_b) { pend = tkB; T[1].tid_out = -1; T[1].retry = 0; }
                    }
                }
            }
        }
        // ---- start the tries that are waiting, once no row is in flight (both halves start at row 0 together) ----
        if (!T[0].run && !T[1].run && (T[0].retry || T[1].retry)) {
#pragma unroll
            for (int X = 0; X < 2; ++X) {
                Half &S = T[X];
                if (!S.retry) continue;
                S.retry = 0; S.run = 1;
                const int qlen = S.qlen, h0 = S.h0;
                EH16[X] = (unsigned short)h0;                // e = 0 in the high byte
                int v = h0 > oe_ins ? h0 - oe_ins : 0;
                if (qlen >= 1) EH16[2 * kT + X] = (unsigned short)v;
                int j = 2;
                for (; j <= qlen && v > P.e_ins; ++j) { v -= P.e_ins; EH16[2 * j * kT + X] = (unsigned short)v; }
                for (; j <= qlen; ++j) EH16[2 * j * kT + X] = 0;
                int best = P.a > -1 ? P.a : -1;
                if (-P.b > best) best = -P.b;
                int w = S.w;
                int lim = (int)((double)(qlen * best + S.end_bonus - P.o_ins) / P.e_ins + 1.);
                lim = lim > 1 ? lim : 1;
                w = w < lim ? w : lim;
                lim = (int)((double)(qlen * best + S.end_bonus - P.o_del) / P.e_del + 1.);
                lim = lim > 1 ? lim : 1;
                w = w < lim ? w : lim;
                S.w = w;
                S.mx = h0; S.mx_i = -1; S.mx_j = -1; S.mx_ie = -1; S.gscore = -1; S.max_off = 0;
                S.beg = 0; S.end = qlen;
                S.tb_next = S.tlen > 0 ? (S.indirect ? qm_ref_base(V, S.t0) : S.t[0]) : 0;
            }
            i = 0;
        }
        if (!T[0].run && !T[1].run) continue;               // waiting for a refill batch
        // ---- one row of both tasks ----
        int rb[2], re[2], h1i[2], tb[2];
        bool act[2];
#pragma unroll
        for (int X = 0; X < 2; ++X) {
            Half &S = T[X];
            act[X] = S.run && i < S.tlen;
            rb[X] = re[X] = 0; h1i[X] = 0; tb[X] = 4;
            if (!act[X]) continue;
            tb[X] = S.tb_next;
            if (i + 1 < S.tlen) S.tb_next = S.indirect ? qm_ref_base(V, S.t0 + (int64_t)(i + 1) * S.tstep) : S.t[i + 1];
            int beg = S.beg, end = S.end;
            if (beg < i - S.w) beg = i - S.w;
            if (end > i + S.w + 1) end = i + S.w + 1;
            if (end > S.qlen) end = S.qlen;
            rb[X] = beg; re[X] = end;
            if (beg == 0) { int h = S.h0 - (P.o_del + P.e_del * (i + 1)); h1i[X] = h > 0 ? h : 0; }
        }
        const unsigned initw = (unsigned)h1i[0] | ((unsigned)h1i[1] << 16);
        unsigned h1fw = initw;                              // per half: h1 after its last cell of this row
        unsigned bkey = 0;                                  // per half: max over the row's cells of H << 7 | column
        if (act[0] || act[1]) {
            const unsigned lutA = make_lut4(P, tb[0]), lutB = make_lut4(P, tb[1]);
            // Up to three column segments per row: the columns only one task visits before the common range, the
            // common range, the columns only one task visits after it (no common column: one segment per task).
            // ONE copy of the loop serves all three with a run-time half-word mask, so that the lanes of a warp --
            // whatever the shape of their two tasks -- walk the same instructions.
            const int a0 = rb[0], a1 = re[0] > rb[0] ? re[0] : rb[0];
            const int b0 = rb[1], b1 = re[1] > rb[1] ? re[1] : rb[1];
            const int ov0 = a0 > b0 ? a0 : b0, ov1 = a1 < b1 ? a1 : b1;
            const unsigned LO = 0x0000ffffu, HI = 0xffff0000u;
            int s0a, s0b, s1a, s1b, s2a, s2b;
            unsigned m0, m1, m2;
            if (ov1 > ov0) {
                s0a = a0 < b0 ? a0 : b0; s0b = ov0; m0 = a0 < b0 ? LO : HI;
                s1a = ov0; s1b = ov1; m1 = LO | HI;
                s2a = ov1; s2b = a1 > b1 ? a1 : b1; m2 = a1 > b1 ? LO : HI;
            } else {
                s0a = s0b = 0; m0 = LO;
                if (a1 > a0) { s1a = a0; s1b = a1; m1 = LO; s2a = b0; s2b = b1; m2 = HI; }
                else { s1a = b0; s1b = b1; m1 = HI; s2a = s2b = 0; m2 = HI; }
            }
            unsigned h1 = 0, f = 0, started = 0;
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                int j = k == 0 ? s0a : k == 1 ? s1a : s2a;
                const int j1 = k == 0 ? s0b : k == 1 ? s1b : s2b;
                const unsigned msk = k == 0 ? m0 : k == 1 ? m1 : m2;
                const bool ran = j < j1;
                if (ran) {
                    const unsigned fresh = msk & ~started;      // halves whose range starts with this segment
                    h1 = (h1 & ~fresh) | (initw & fresh);
                    f &= ~fresh;
                    started |= msk;
                }
                // one cell at word offset O from hw, selector SV, column offset inside the group KC; MASKED: blend the
                // store and the key with the segment's half-word mask (otherwise both tasks are live)
#define QM_CELL2(O, SV, KC, K, MASKED)                                                             \
                {                                                                                  \
                    const unsigned w = hw[O];                                                      \
                    const unsigned hh = w & 0x00ff00ffu, e = prmt(w, 0u, 0x4341u);                 \
                    const unsigned s = prmt(lutA, lutB, SV);                                       \
                    const unsigned M = add2(hh, min2(s, hh));                                      \
                    const unsigned H = __vimax3_s16x2(M, e, f);                                    \
                    const unsigned td = add2(M, noe_del2);                                         \
                    const unsigned en = __viaddmax_s16x2_relu(e, ned2, td);                        \
                    f = __viaddmax_s16x2_relu(f, nei2, SYM ? td : add2(M, noe_ins2));              \
                    const unsigned nw = en * 256u + h1;                                            \
                    hw[O] = (MASKED) ? (nw & msk) | (w & ~msk) : nw;                               \
                    h1 = H;                                                                        \
                    K = ((MASKED) ? H & msk : H) * 128u + (KC);                                    \
                }
#define QM_ROW_LOOPS(MASKED)                                                                       \
                if ((j & 1) && j < j1) {                                                           \
                    unsigned *hw = HW + j * kT;                                                    \
                    const unsigned sv = SW[(j >> 1) * kT] >> 16;                                   \
                    unsigned k0;                                                                   \
                    QM_CELL2(0, sv, 0u, k0, MASKED)                                                \
                    bkey = __viaddmax_u16x2(k0, pack2(j), bkey);                                   \
                    ++j;                                                                           \
                }                                                                                  \
                for (; j + 3 < j1; j += 4) {                                                       \
                    unsigned *hw = HW + j * kT;                                                    \
                    const unsigned s01 = SW[(j >> 1) * kT], s23 = SW[((j >> 1) + 1) * kT];         \
                    unsigned k0, k1, k2, k3;                                                       \
                    QM_CELL2(0, s01, 0u, k0, MASKED) QM_CELL2(kT, s01 >> 16, kc1, k1, MASKED)      \
                    QM_CELL2(2 * kT, s23, kc2, k2, MASKED) QM_CELL2(3 * kT, s23 >> 16, kc3, k3, MASKED) \
                    const unsigned m4 = maxu2(__vimax3_u16x2(k0, k1, k2), k3);                     \
                    bkey = __viaddmax_u16x2(m4, pack2(j), bkey);                                   \
                }                                                                                  \
                for (; j < j1; ++j) {                                                              \
                    unsigned *hw = HW + j * kT;                                                    \
                    const unsigned sv = SW[(j >> 1) * kT] >> ((j & 1) * 16);                       \
                    unsigned k0;                                                                   \
                    QM_CELL2(0, sv, 0u, k0, MASKED)                                                \
                    bkey = __viaddmax_u16x2(k0, pack2(j), bkey);                                   \
                }
                // warp-uniform choice: when every lane here has both tasks live in this segment (or nothing to do),
                // the stores and keys need no blend
                if (__all_sync(__activemask(), !ran || msk == 0xffffffffu)) { QM_ROW_LOOPS(false) }
                else { QM_ROW_LOOPS(true) }
#undef QM_ROW_LOOPS
#undef QM_CELL2
                if (ran) h1fw = (h1fw & ~msk) | (h1 & msk);
            }
        }
        const int h1f[2] = { (int)(h1fw & 0xffffu), (int)(h1fw >> 16) };
        // ---- per task: end of the row, then end of the try ----
#pragma unroll
        for (int X = 0; X < 2; ++X) {
            Half &S = T[X];
            if (!S.run) continue;
            bool done = true;
            if (act[X]) {
                done = false;
                const int beg = rb[X], end = re[X], qlen = S.qlen, h1 = h1f[X];
                const int jstop = end > beg ? end : beg;
                EH16[2 * jstop * kT + X] = (unsigned short)h1;
                if (end > beg) S.cells += end - beg;
                int m = 0, mj = -1;
                if (end > beg) { const unsigned k = X ? bkey >> 16 : bkey & 0xffffu; m = (int)(k >> 7); mj = (int)(k & 127u); }
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
                    int a = beg;
                    while (a < end && EH16[2 * a * kT + X] == 0) ++a;
                    int b = end;
                    while (b >= a && EH16[2 * b * kT + X] == 0) --b;
                    S.beg = a;
                    S.end = b + 2 < qlen ? b + 2 : qlen;
                    if (i + 1 >= S.tlen) done = true;
                }
            }
            if (done) {
                S.run = 0;
                const int wu = S.w0;
                if (S.tries_left == 2 && !(S.mx == S.prev || S.max_off < (wu >> 1) + (wu >> 2))) {
                    S.prev = S.mx; S.tries_left = 1; S.w0 = wu << 1; S.w = S.w0; S.retry = 1;
                } else {
                    qm_ext_result o;
                    o.score = S.mx; o.qle = S.mx_j + 1; o.tle = S.mx_i + 1; o.gtle = S.mx_ie + 1; o.gscore = S.gscore;
                    o.max_off = S.max_off; o.w_used = wu; o.cells = S.cells;
                    out[S.tid_out] = o;
                    S.tid_out = -1;
                }
            }
        }
        ++i;
    }
}

template <int CAP>
void launch2p(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks, const int *d_lists,
              int64_t list_stride, const int *d_counts, int *d_cursors, int h_count, qm_ext_result *d_out,
              int *d_fb_lists, int *d_fb_ctr, cudaStream_t st)
{
    const size_t smem = ((size_t)(CAP + 1) + (CAP + 2) / 2) * kT * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(ext2p_kernel<CAP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(ext2p_kernel<CAP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    int per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 16) per_sm = 16;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (h_count >= 0) {
        const int64_t need = ((h_count + 1) / 2 + kT - 1) / kT;
        if (need < blocks) blocks = need;
    }
    if (blocks < 1) return;
    const bool sym = P.o_del == P.o_ins && P.e_del == P.e_ins;
    if (sym) ext2p_kernel<CAP, true><<<(unsigned)blocks, kT, smem, st>>>(P, V, d_tasks, d_lists + cls * list_stride, d_counts + cls, d_cursors + cls, d_out, d_fb_lists + cls * list_stride, d_fb_ctr + cls);
    else ext2p_kernel<CAP, false><<<(unsigned)blocks, kT, smem, st>>>(P, V, d_tasks, d_lists + cls * list_stride, d_counts + cls, d_cursors + cls, d_out, d_fb_lists + cls * list_stride, d_fb_ctr + cls);
}

}  // namespace

// classes 0..7 (qlen <= 128; a query of exactly 128 bases goes to the fallback list)
int qm_ext2p_launch_class(qm_ctx *ctx, int cls, const ExtParams &P, const IndexView &V, const ExtTaskI *d_tasks,
                          const int *d_lists, int64_t list_stride, const int *d_counts, int *d_cursors, int h_count,
                          qm_ext_result *d_out, int *d_fb_lists, int *d_fb_ctr, cudaStream_t st)
{
    switch (cls) {
    case 0: launch2p<16>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 1: launch2p<32>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 2: launch2p<48>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 3: launch2p<64>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 4: launch2p<80>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 5: launch2p<96>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 6: launch2p<112>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    case 7: launch2p<128>(ctx, cls, P, V, d_tasks, d_lists, list_stride, d_counts, d_cursors, h_count, d_out, d_fb_lists, d_fb_ctr, st); break;
    default: return qm_fail(ctx, QM_EINVAL, "qm_ext2p_launch_class: class %d has no paired kernel", cls);
    }
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}
