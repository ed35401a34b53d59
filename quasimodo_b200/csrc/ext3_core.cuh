// ext3_core.cuh -- batched ksw_extend2, formulation "P2": ONE THREAD PER PAIR OF TASKS, packed s16x2 cells.
// (bwa 0.7.17 ksw.c:ksw_extend2 as called by bwamem.c:mem_chain2aln; reference call site rules/bwa.smk:15;
// semantics SURVEY.md A.3.)
//
// Two independent tasks A and B share a thread: A lives in the low, B in the high 16 bits of every DP word.  The tasks
// of a launch come sorted by query length, so the two tasks of a thread have the same number of columns and (the target
// window being query length + the longest gap the scores allow) almost always the same number of rows; their column
// ranges [beg, end) coincide on most rows.  A row of both tasks is then ONE loop over the common columns in which every
// DPX instruction updates a cell of each task:
//     an = q2 & tmask2                      query base one-hot (per column) against the row's target base
//     s2 = min(an + c2, a2)                 VIADDMNMX.S16x2   match a / mismatch -b / N -1   (see e3_qcode)
//     M2 = h2 + min(s2, h2)                 VIMNMX.S16x2, VIADD.16x2   ("M = h ? h + s : 0", a dead diagonal stays dead)
//     H2 = max3(M2, e2, f2)                 VIMNMX3.S16x2
//     td = max(M2 - oe, 0)                  VIADDMNMX.S16x2.RELU
//     e2' = max(e2 - e_del, td)             VIADDMNMX.S16x2
//     f2' = max(f2 - e_ins, td)             VIADDMNMX.S16x2
// plus the row maximum as a packed unsigned key H << 8 | column (ties go to the later column, like the reference's
// "mj = m > h ? mj : j"), tracked once per four columns.  About 9.5 ALU-pipe instructions per PAIR of cells, against the
// 9 per two cells the roofline accounting assumes.  Columns only one of the two tasks visits (its range starts earlier or
// ends later than the partner's) run the same cell under a half-word blend, so the other task's cells keep their (stale)
// values exactly as the reference's in-place eh[] array does.  Each half keeps its own row counter, band, z-drop state and
// band retry: the two tasks are independent in everything but the shared instruction stream.
//
// Shared memory per thread and column: one 64-bit word {h2, e2} = the reference's eh[j] of both tasks, and one 32-bit word
// with the two query codes (kE3Pad extra columns behind the query plane: the row loop fetches one trip ahead).  Layout [column][thread]: a warp's accesses hit consecutive banks whatever column each lane is at.
//
// Limits (the launcher sends everything else to the scalar kernel of extend2.cu): scores <= 255 and columns <= 256 (the
// packed key), 1 <= a <= 127, 1 <= b, a + b <= 256, gap penalties < 2^14.
//
// The per-thread logic below is plain C++ over a memory accessor, compiled for the device by extend3.cu and for the HOST by
// tests/ext3_host.cpp, where the DPX instructions are emulated: the CPU test suite runs the very statements the kernel
// runs against the oracle, without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define E3_HD __host__ __device__ __forceinline__
#else
#define E3_HD inline
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
#endif

// ---- packed 16-bit pair arithmetic: DPX on the device, emulated on the host ----
#if defined(__CUDA_ARCH__)
E3_HD unsigned e3_addmin(unsigned a, unsigned b, unsigned c) { return __viaddmin_s16x2(a, b, c); }       // min(a + b, c)
E3_HD unsigned e3_addmax(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2(a, b, c); }       // max(a + b, c)
E3_HD unsigned e3_addmax_relu(unsigned a, unsigned b, unsigned c) { return __viaddmax_s16x2_relu(a, b, c); }
E3_HD unsigned e3_max3(unsigned a, unsigned b, unsigned c) { return __vimax3_s16x2(a, b, c); }
E3_HD unsigned e3_min(unsigned a, unsigned b) { return __vmins2(a, b); }
E3_HD unsigned e3_add(unsigned a, unsigned b) { return __vadd2(a, b); }
E3_HD unsigned e3_umax(unsigned a, unsigned b) { return __vmaxu2(a, b); }
E3_HD unsigned e3_umax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
E3_HD unsigned e3_uaddmax(unsigned a, unsigned b, unsigned c) { return __viaddmax_u16x2(a, b, c); }      // unsigned max(a + b, c)
#else
namespace e3emu {
inline int lo(unsigned x) { return (int16_t)(x & 0xffffu); }
inline int hi(unsigned x) { return (int16_t)(x >> 16); }
inline unsigned pk(int l, int h) { return ((unsigned)l & 0xffffu) | ((unsigned)h << 16); }
inline int w16(int v) { return (int16_t)(uint16_t)v; }                      // wrap to 16 bits like the hardware
inline int mx(int a, int b) { return a > b ? a : b; }
inline int mn(int a, int b) { return a < b ? a : b; }
inline unsigned ulo(unsigned x) { return x & 0xffffu; }
inline unsigned uhi(unsigned x) { return x >> 16; }
inline unsigned umx(unsigned a, unsigned b) { return a > b ? a : b; }
}
E3_HD unsigned e3_addmin(unsigned a, unsigned b, unsigned c) { using namespace e3emu; return pk(mn(w16(lo(a) + lo(b)), lo(c)), mn(w16(hi(a) + hi(b)), hi(c))); }
E3_HD unsigned e3_addmax(unsigned a, unsigned b, unsigned c) { using namespace e3emu; return pk(mx(w16(lo(a) + lo(b)), lo(c)), mx(w16(hi(a) + hi(b)), hi(c))); }
E3_HD unsigned e3_addmax_relu(unsigned a, unsigned b, unsigned c) { using namespace e3emu; return pk(mx(mx(w16(lo(a) + lo(b)), lo(c)), 0), mx(mx(w16(hi(a) + hi(b)), hi(c)), 0)); }
E3_HD unsigned e3_max3(unsigned a, unsigned b, unsigned c) { using namespace e3emu; return pk(mx(mx(lo(a), lo(b)), lo(c)), mx(mx(hi(a), hi(b)), hi(c))); }
E3_HD unsigned e3_min(unsigned a, unsigned b) { using namespace e3emu; return pk(mn(lo(a), lo(b)), mn(hi(a), hi(b))); }
E3_HD unsigned e3_add(unsigned a, unsigned b) { using namespace e3emu; return pk(lo(a) + lo(b), hi(a) + hi(b)); }
E3_HD unsigned e3_umax(unsigned a, unsigned b) { using namespace e3emu; return umx(ulo(a), ulo(b)) | umx(uhi(a), uhi(b)) << 16; }
E3_HD unsigned e3_umax3(unsigned a, unsigned b, unsigned c) { return e3_umax(e3_umax(a, b), c); }
E3_HD unsigned e3_uaddmax(unsigned a, unsigned b, unsigned c) { using namespace e3emu; return umx((ulo(a) + ulo(b)) & 0xffffu, ulo(c)) | umx((uhi(a) + uhi(b)) & 0xffffu, uhi(c)) << 16; }
#endif

// host-side instrumentation of the test build (tests/ext3_host.cpp -DE3_STATS): where the columns go
#if defined(E3_STATS) && !defined(__CUDA_ARCH__)
struct E3Stats { long long rows, core_cols, masked_cols, solo_rows; int last_pre, last_core, last_post; };
extern E3Stats g_e3_stats;
#define E3_COUNT(field, n) (g_e3_stats.field += (n))
#else
#define E3_COUNT(field, n) ((void)0)
#endif

E3_HD unsigned e3_pack2(int v) { return ((unsigned)v & 0xffffu) * 0x10001u; }
constexpr int kE3Pad = 4;                         // columns the one-trip-ahead fetch may read past a range
constexpr unsigned kE3Neg = 0x80008000u;         // -32768 | -32768: the neutral third operand of max(a + b, c)

struct E3Scores { int a, b, o_del, e_del, o_ins, e_ins, zdrop; };

// can the packed kernel run this scoring scheme at all?
E3_HD bool e3_scores_ok(const E3Scores &S)
{
    return S.a >= 1 && S.a <= 127 && S.b >= 1 && S.a + S.b <= 256 && S.o_del >= 0 && S.o_ins >= 0 && S.e_del >= 1 && S.e_ins >= 1 &&
           S.o_del + S.e_del < 16384 && S.o_ins + S.e_ins < 16384;
}
// can it hold this task?  (packed key: score <= 255 in 8 bits, column <= 255 in 8 bits)
E3_HD bool e3_task_ok(const E3Scores &S, int qlen, int h0, int cap) { return qlen <= cap && qlen <= 256 && h0 + qlen * S.a <= 255 && h0 >= 0; }

// Substitution score without a table: the query base is stored per column as a small code, the row's target base becomes a
// mask and a floor:   s = min((qcode & tmask) + floor, a)          (U = 0x100 with 16-bit planes, 0x10 with byte planes)
//   query A/C/G/T : U << base                query N : b - 1 (below U)
//   target A/C/G/T: mask U << base | (U - 1), floor -b         target N: mask 0, floor -1
// match: >= U - b >= a -> a.  mismatch: 0 - b.  query N vs base: (b - 1) - b = -1.  anything vs target N: 0 - 1 = -1.
template <bool NARROW> E3_HD unsigned e3_qcode(const E3Scores &S, int c) { return c > 3 ? (unsigned)(S.b - 1) & (NARROW ? 0xfu : 0xffu) : (NARROW ? 0x10u : 0x100u) << c; }
template <bool NARROW> E3_HD unsigned e3_tmask(int tb) { return tb > 3 ? 0u : ((NARROW ? 0x10u : 0x100u) << tb) | (NARROW ? 0xfu : 0xffu); }
E3_HD unsigned e3_tfloor(const E3Scores &S, int tb) { return (unsigned)(tb > 3 ? -1 : -S.b) & 0xffffu; }
// byte planes: the query code is a byte, so the scores must satisfy a + b <= 16 (the default scheme does)
E3_HD bool e3_scores_ok_narrow(const E3Scores &S) { return e3_scores_ok(S) && S.a + S.b <= 16; }

struct E3Consts {
    E3Scores S;
    unsigned a2, noe_del2, noe_ins2, ned2, nei2;
    int oe_del, oe_ins;
};
E3_HD E3Consts e3_consts(const E3Scores &S)
{
    E3Consts K;
    K.S = S;
    K.oe_del = S.o_del + S.e_del; K.oe_ins = S.o_ins + S.e_ins;
    K.a2 = e3_pack2(S.a); K.noe_del2 = e3_pack2(-K.oe_del); K.noe_ins2 = e3_pack2(-K.oe_ins);
    K.ned2 = e3_pack2(-S.e_del); K.nei2 = e3_pack2(-S.e_ins);
    return K;
}

// one of the two tasks of a thread (registers)
struct E3Half {
    int tk;                 // result slot, -1 = no task
    int phase;              // 0 = none, 1 = a try is about to start (row -1 not written yet), 2 = rows in progress
    int qlen, tlen, h0, w0, w, end_bonus, tries_left, prev, cells;
    int i, beg, end;
    int mx, mx_i, mx_j, mx_ie, gscore, max_off;
    int tb_next;            // target byte of row i as fetched one row ahead (Tgt::decode turns it into a base code when the row
                            // starts: nothing may touch the loaded value earlier, or the row waits for the load at once)
};

struct E3Result { int score, qle, tle, gtle, gscore, max_off, w_used, cells; };

// Mem (the thread's columns; two layouts, see extend3.cu):
//   Raw raw(j)                      the stored words of column j, as loaded;  unpack(raw, h2, e2, q2) -> packed halves
//   put(j, h2, e2)                  store eh[j] of both tasks
//   set_he(j, X, h, e), zero(j, X)  eh[j] of task X alone (row -1, eh[end], the trimming scans)
//   set_q(j, X, code)               query code of task X at column j;  Mem::kNarrow tells which code set is in use
// Tgt: raw(X, i) -> the stored byte of row i of task X's target; decode(X, byte) -> its base code 0..4.

// row -1 of one half (SURVEY.md A.3 first lines) and the band clamp of this try
template <class Mem, class Tgt>
E3_HD void e3_start_try(const E3Consts &K, E3Half &H, int X, Mem &mem, Tgt &tgt)
{
    const E3Scores &S = K.S;
    const int qlen = H.qlen, h0 = H.h0;
    mem.set_he(0, X, h0, 0);
    int v = h0 > K.oe_ins ? h0 - K.oe_ins : 0;
    if (qlen >= 1) mem.set_he(1, X, v, 0);
    int j = 2;
    for (; j <= qlen && v > S.e_ins; ++j) { v -= S.e_ins; mem.set_he(j, X, v, 0); }
    for (; j <= qlen; ++j) mem.set_he(j, X, 0, 0);
    int best = S.a > -1 ? S.a : -1;
    if (-S.b > best) best = -S.b;
    int w = H.w0;
    int lim = (int)((double)(qlen * best + H.end_bonus - S.o_ins) / S.e_ins + 1.);
    lim = lim > 1 ? lim : 1;
    w = w < lim ? w : lim;
    lim = (int)((double)(qlen * best + H.end_bonus - S.o_del) / S.e_del + 1.);
    lim = lim > 1 ? lim : 1;
    w = w < lim ? w : lim;
    H.w = w;                                         // clamped band of this try (w_used reports the unclamped one)
    H.mx = h0; H.mx_i = -1; H.mx_j = -1; H.mx_ie = -1; H.gscore = -1; H.max_off = 0;
    H.beg = 0; H.end = qlen; H.i = 0;
    H.tb_next = H.tlen > 0 ? tgt.raw(X, 0) : 0;
    H.phase = 2;
}

// the query codes of a freshly loaded task into half X of the query plane.  Qry: code(j) -> base code 0..4 of column j
template <class Mem, class Qry>
E3_HD void e3_load_query(const E3Consts &K, int qlen, int X, Mem &mem, Qry &qry)
{
    // eight bases per trip, all eight fetched before the first is stored: the fetches are global byte loads, and one at a time
    // they were 8 % of the kernel's warp time at 0.4 % of its instructions
    int j = 0;
    for (; j + 8 <= qlen; j += 8) {
        int c[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 8; ++u) c[u] = qry.code(j + u);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int u = 0; u < 8; ++u) mem.set_q(j + u, X, e3_qcode<Mem::kNarrow>(K.S, c[u]));
    }
    for (; j < qlen; ++j) mem.set_q(j, X, e3_qcode<Mem::kNarrow>(K.S, qry.code(j)));
}

// the end of a try of half H: second try with a doubled band (mem_chain2aln's rule), or the result.  Returns true when
// the task is finished and *out holds its result.
E3_HD bool e3_end_try(E3Half &H, E3Result *out)
{
    const int wu = H.w0;                             // unclamped band of this try (what w_used reports)
    if (H.tries_left == 2 && !(H.mx == H.prev || H.max_off < (wu >> 1) + (wu >> 2))) {
        H.prev = H.mx; H.tries_left = 1; H.w0 = wu << 1; H.phase = 1;
        return false;
    }
    out->score = H.mx; out->qle = H.mx_j + 1; out->tle = H.mx_i + 1; out->gtle = H.mx_ie + 1; out->gscore = H.gscore;
    out->max_off = H.max_off; out->w_used = wu; out->cells = H.cells;
    H.phase = 0;
    return true;
}

// One row of both halves.  doneA / doneB: that half's try ended with this call (no more rows, m == 0 or z-drop).
template <bool SYM, class Mem, class Tgt>
E3_HD void e3_row(const E3Consts &K, E3Half &A, E3Half &B, Mem &mem, Tgt &tgt, bool &doneA, bool &doneB)
{
    const E3Scores &S = K.S;
    const bool ra = A.phase == 2 && A.i < A.tlen, rb = B.phase == 2 && B.i < B.tlen;
    doneA = A.phase == 2 && !ra; doneB = B.phase == 2 && !rb;
    if (!ra && !rb) return;
    // ---- per-half row set-up: band clamp, first-column score, the row's target base as mask + floor ----
    int begA = 0, endA = 0, begB = 0, endB = 0, h1A = 0, h1B = 0;
    unsigned tmA = 0, tmB = 0, flA = 0, flB = 0;
    if (ra) {
        const int i = A.i, tb = tgt.decode(0, A.tb_next);
        if (i + 1 < A.tlen) A.tb_next = tgt.raw(0, i + 1);
        begA = A.beg > i - A.w ? A.beg : i - A.w;
        endA = A.end < i + A.w + 1 ? A.end : i + A.w + 1;
        endA = endA < A.qlen ? endA : A.qlen;
        if (begA == 0) { h1A = A.h0 - (S.o_del + S.e_del * (i + 1)); h1A = h1A > 0 ? h1A : 0; }
        tmA = e3_tmask<Mem::kNarrow>(tb); flA = e3_tfloor(S, tb);
    }
    if (rb) {
        const int i = B.i, tb = tgt.decode(1, B.tb_next);
        if (i + 1 < B.tlen) B.tb_next = tgt.raw(1, i + 1);
        begB = B.beg > i - B.w ? B.beg : i - B.w;
        endB = B.end < i + B.w + 1 ? B.end : i + B.w + 1;
        endB = endB < B.qlen ? endB : B.qlen;
        if (begB == 0) { h1B = B.h0 - (S.o_del + S.e_del * (i + 1)); h1B = h1B > 0 ? h1B : 0; }
        tmB = e3_tmask<Mem::kNarrow>(tb); flB = e3_tfloor(S, tb);
    }
    // a half without a row rides along on the partner's columns: its cells are dead storage (a finished task, or one whose
    // next try rewrites row -1 first), so the common loop needs no mask for it
    int cbA = begA, ceA = endA, cbB = begB, ceB = endB;
    if (!ra) { cbA = begB; ceA = endB; }
    if (!rb) { cbB = begA; ceB = endA; }
    const unsigned tm2 = tmA | tmB << 16, fl2 = flA | flB << 16;
    unsigned h1 = ((unsigned)h1A & 0xffffu) | (unsigned)h1B << 16;
    unsigned f2 = 0, bkey = 0;

    // one cell of each task at column j: the reference's inner loop body on packed halves; KEY = H << 8 (both halves)
#define E3_CELL(H2, E2, Q2, HOUT, EOUT)                                                                               \
    {                                                                                                                 \
        const unsigned an = (Q2) & tm2;                                                                               \
        const unsigned s2 = e3_addmin(an, fl2, K.a2);                                                                 \
        const unsigned m2 = e3_min(s2, SYM ? (H2) : (H2) * 128u);          /* h ? s : min(s, 0): a dead diagonal stays dead */ \
        const unsigned M2 = e3_add((H2), m2);                                                                         \
        HOUT = e3_max3(M2, (E2), f2);                                                                                 \
        const unsigned td = e3_addmax_relu(M2, K.noe_del2, kE3Neg);                                                   \
        EOUT = e3_addmax((E2), K.ned2, td);                                                                           \
        f2 = e3_addmax(f2, K.nei2, SYM ? td : e3_addmax_relu(M2, K.noe_ins2, kE3Neg));                                \
    }
    // ---- columns only one task visits before the common range ----
    // Left of its beg a task only has cells the zero-span trimming dropped (h = e = 0, untouched since): running the cell on
    // them writes the same zeros back, so the task whose range starts later simply starts with its partner -- as long as the
    // columns are inside its band (a column the BAND cut off keeps a non-zero stale value and must stay untouched).  The
    // executed-cell count and the trimming still use the task's own beg.  The right edge has no such freedom: past end the
    // reference drops insertion tails that the cell would carry on.
    int cs = cbA < cbB ? cbA : cbB;
    if (ra) { const int z = A.i - A.w; cs = cs > z ? cs : z; }
    if (rb) { const int z = B.i - B.w; cs = cs > z ? cs : z; }
    int ce = ceA < ceB ? ceA : ceB;
    ce = ce > cs ? ce : cs;
    auto masked = [&](int j0, int j1, unsigned mask) {
        E3_COUNT(masked_cols, j1 > j0 ? j1 - j0 : 0);
        E3_COUNT(last_pre, j1 > j0 ? j1 - j0 : 0);            /* columns before + after the common range, this row */
        if (j0 >= j1) return;
        typename Mem::Raw cur = mem.raw(j0);
        for (int j = j0; j < j1; ++j) {
            // (the next column is fetched before this one's arithmetic: with two or three lanes in here the loop is a latency chain)
            const typename Mem::Raw nxt = mem.raw(j + 1 < j1 ? j + 1 : j);
            unsigned h2, e2, q2, Hn, En;
            Mem::unpack(cur, h2, e2, q2);
            cur = nxt;
            const unsigned fkeep = f2;
            E3_CELL(h2, e2, q2, Hn, En)
            mem.put(j, (h1 & mask) | (h2 & ~mask), (En & mask) | (e2 & ~mask));
            Hn &= mask;                      // the other half computed on cells it does not own: unbounded garbage, keep it out of the key
            h1 = Hn | (h1 & ~mask);
            f2 = (f2 & mask) | (fkeep & ~mask);
            bkey = e3_umax(bkey, Hn * 256u + (e3_pack2(j) & mask));
        }
    };
    if (ra && rb) {
        if (begA < cs) masked(begA, endA < cs ? endA : cs, 0x0000ffffu);
        else if (begB < cs) masked(begB, endB < cs ? endB : cs, 0xffff0000u);
    }
    // ---- the common range: both tasks per instruction, four columns per trip.  The next trip's eight shared-memory words are
    // fetched before this trip's arithmetic starts: a thread's columns are private, so nothing written here can change them,
    // and with only a few warps per scheduler the loads have to be in flight early.  (The fetch may run up to four columns
    // past the range: the planes are padded, the values unused.) ----
    {
        E3_COUNT(rows, 1); E3_COUNT(core_cols, ce - cs); E3_COUNT(solo_rows, (ra && rb) ? 0 : 1); E3_COUNT(last_core, ce - cs);
        int j = cs;
        unsigned jb = e3_pack2(cs);
#define E3_STEP(JJ, RAW, U, KU)                                                              \
            {                                                                                \
                unsigned h2, e2, q2, Hn, En;                                                 \
                Mem::unpack(RAW, h2, e2, q2);                                                \
                E3_CELL(h2, e2, q2, Hn, En)                                                  \
                mem.put(JJ, h1, En);                                                         \
                h1 = Hn;                                                                     \
                KU = Hn * 256u + (unsigned)((U) * 0x10001);                                  \
            }
        if (j + 3 < ce) {
            typename Mem::Raw a0 = mem.raw(j), a1 = mem.raw(j + 1), a2 = mem.raw(j + 2), a3 = mem.raw(j + 3);
            do {
                const typename Mem::Raw n0 = mem.raw(j + 4), n1 = mem.raw(j + 5), n2 = mem.raw(j + 6), n3 = mem.raw(j + 7);
                unsigned k0, k1, k2, k3;
                E3_STEP(j, a0, 0, k0) E3_STEP(j + 1, a1, 1, k1) E3_STEP(j + 2, a2, 2, k2) E3_STEP(j + 3, a3, 3, k3)
                bkey = e3_uaddmax(e3_umax(e3_umax3(k0, k1, k2), k3), jb, bkey);
                jb += 0x00040004u;
                j += 4;
                a0 = n0; a1 = n1; a2 = n2; a3 = n3;
            } while (j + 3 < ce);
            // up to three columns left, already in registers
            if (j < ce) { unsigned k0; E3_STEP(j, a0, 0, k0) bkey = e3_uaddmax(k0, jb, bkey); }
            if (j + 1 < ce) { unsigned k0; E3_STEP(j + 1, a1, 1, k0) bkey = e3_uaddmax(k0, jb, bkey); }
            if (j + 2 < ce) { unsigned k0; E3_STEP(j + 2, a2, 2, k0) bkey = e3_uaddmax(k0, jb, bkey); }
        } else {
            for (; j < ce; ++j) {
                const typename Mem::Raw r0 = mem.raw(j);
                unsigned k0;
                E3_STEP(j, r0, 0, k0)
                bkey = e3_uaddmax(k0, jb, bkey);
                jb += 0x00010001u;
            }
        }
#undef E3_STEP
    }
    // ---- columns only one task visits after the common range ----
    if (ra && rb) {
        if (endA > ce) masked(begA > ce ? begA : ce, endA, 0x0000ffffu);
        else if (endB > ce) masked(begB > ce ? begB : ce, endB, 0xffff0000u);
    }
#undef E3_CELL
    // ---- per-half row end: eh[end], to-end score, row maximum, z-drop, zero-span trimming ----
    auto row_end = [&](E3Half &H, int X, int beg, int end, bool &done) {
        const int i = H.i, qlen = H.qlen;
        const int h1x = (int)((h1 >> (16 * X)) & 0xffffu);
        const unsigned key = (bkey >> (16 * X)) & 0xffffu;
        const int jstop = end > beg ? end : beg;                      // the reference's j after the loop
        mem.set_he(end, X, h1x, 0);                                   // eh[end] = {h1, 0}
        int m = 0, mj = -1;
        if (end > beg) { H.cells += end - beg; m = (int)(key >> 8); mj = (int)(key & 0xffu); }
        if (jstop == qlen) {
            H.mx_ie = H.gscore > h1x ? H.mx_ie : i;
            H.gscore = H.gscore > h1x ? H.gscore : h1x;
        }
        bool stop = m == 0;
        if (!stop) {
            if (m > H.mx) {
                H.mx = m; H.mx_i = i; H.mx_j = mj;
                const int d = mj > i ? mj - i : i - mj;
                H.max_off = H.max_off > d ? H.max_off : d;
            } else if (S.zdrop > 0) {
                const int dr = i - H.mx_i, dc = mj - H.mx_j;
                if (dr > dc) { if (H.mx - m - (dr - dc) * S.e_del > S.zdrop) stop = true; }
                else         { if (H.mx - m - (dc - dr) * S.e_ins > S.zdrop) stop = true; }
            }
        }
        if (stop) { done = true; return; }
        // zero-span trimming ("for (j = beg; j < end && eh[j] == 0; ++j); beg = j; for (j = end; j >= beg && eh[j] == 0; --j);"), four
        // columns per trip: the four loads are independent, so a trip costs ONE shared-memory latency instead of four (the scans
        // are latency, not instructions: 13 % of the kernel's warp time at 4 % of its instructions).  Columns outside the span are
        // read at a clamped index and not counted.
        int a = beg;
        while (a < end) {
            const int lim = end - a;
            const bool z0 = mem.zero(a, X), z1 = mem.zero(a + 1 < end ? a + 1 : end, X), z2 = mem.zero(a + 2 < end ? a + 2 : end, X),
                       z3 = mem.zero(a + 3 < end ? a + 3 : end, X);
            int z = !z0 ? 0 : !z1 ? 1 : !z2 ? 2 : !z3 ? 3 : 4;
            z = z < lim ? z : lim;
            a += z;
            if (z < 4) break;
        }
        int b = end;
        while (b >= a) {
            const int lim = b - a + 1;
            const bool z0 = mem.zero(b, X), z1 = mem.zero(b - 1 > 0 ? b - 1 : 0, X), z2 = mem.zero(b - 2 > 0 ? b - 2 : 0, X),
                       z3 = mem.zero(b - 3 > 0 ? b - 3 : 0, X);
            int z = !z0 ? 0 : !z1 ? 1 : !z2 ? 2 : !z3 ? 3 : 4;
            z = z < lim ? z : lim;
            b -= z;
            if (z < 4) break;
        }
        H.beg = a;
        H.end = b + 2 < qlen ? b + 2 : qlen;
        H.i = i + 1;
        if (H.i >= H.tlen) done = true;
    };
    if (ra) row_end(A, 0, begA, endA, doneA);
    if (rb) row_end(B, 1, begB, endB, doneB);
}
