// pair.cu -- paired-end stage: insert-size model, primary marking, pairing, MAPQ, CIGAR/NM, SAM flags.
// Replaces the compute of `bwa mem` worker2 (bwamem_pair.c mem_pestat, mem_pair, mem_sam_pe without mate
// rescue; bwamem.c mem_mark_primary_se, mem_approx_mapq_se, mem_reg2aln; bwa.c bwa_gen_cigar2;
// ksw.c ksw_global2 -- reference call site rules/bwa.smk:15; semantics SURVEY.md A.4-A.6).
// One thread per pair.  The floating-point pieces (log / erfc) are evaluated on the HOST into small
// tables (same libm as the CPU path) so that the integer MAPQ / pair scores are bit-identical; the
// device only does IEEE add/mul/div on those table values (the library is built with -fmad=false).
#include <math.h>
#include <algorithm>
#include <vector>
#include "pipeline.cuh"

namespace {

constexpr int kMapqTabLen = 1024;
constexpr int kSubnTabLen = 256;
constexpr int kSliceBytes = 24 * 1024;       // per-thread DP scratch: eh[] (4 KB) + direction matrix
constexpr int kEhBytes = 4 * 1024;
constexpr size_t kOverflowBytes = 256u << 20;

struct PairTables {
    const double *mapq_l;        // [kMapqTabLen]: l < coef_len ? 1 : log(coef_len)/log(l)
    const int *subn;             // [kSubnTabLen]: (int)(4.343*log(n+1)+.499)
    const double *pair_term[4];  // per orientation: .721*log(2*erfc(|ns|/sqrt2))*a for dist = low..high
    qm_pestat pes[4];
};

struct PairScratch {
    uint8_t *slices;             // [threads][kSliceBytes]
    uint8_t *overflow;           // bump pool for large direction matrices
    unsigned long long *overflow_used;
    int *err;
};

__device__ __forceinline__ uint64_t hash64(uint64_t key)
{
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3);   key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

__device__ __forceinline__ int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
    const int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
    const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
    *dist = p2 > b1 ? p2 - b1 : b1 - p2;
    return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

__device__ int cal_sub(const qm_opt &o, const qm_reg *a, int n)
{
    int j;
    for (j = 1; j < n; ++j) {
        const int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
        const int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
        if (e_min > b_max) {
            const int l0 = a[0].qe - a[0].qb, lj = a[j].qe - a[j].qb;
            const int min_l = lj < l0 ? lj : l0;
            if (e_min - b_max >= min_l * o.mask_level) break;
        }
    }
    return j < n ? a[j].score : o.min_seed_len * o.a;
}

// insert-size histogram for mem_pestat: hist[dir][isize] += 1 for every qualifying pair
__global__ void pestat_hist_kernel(IndexView V, qm_opt o, const qm_reg *__restrict__ regs, const int32_t *__restrict__ n_regs,
                                   int64_t n_pairs, unsigned *__restrict__ hist)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const qm_reg *r0 = regs + (2 * i) * QM_MAX_REGS, *r1 = regs + (2 * i + 1) * QM_MAX_REGS;
    const int n0 = n_regs[2 * i], n1 = n_regs[2 * i + 1];
    if (n0 && n1 && !(cal_sub(o, r0, n0) > 0.8 * r0[0].score) && !(cal_sub(o, r1, n1) > 0.8 * r1[0].score) &&
        r0[0].rid == r1[0].rid) {
        int64_t is;
        const int dir = infer_dir(V.l_pac, r0[0].rb, r1[0].rb, &is);
        if (is && is <= o.max_ins) atomicAdd(&hist[(int64_t)dir * (o.max_ins + 1) + is], 1u);
    }
}

// mem_mark_primary_se: sort by (score desc, hash asc), mark secondaries, fill sub / sub_n
__device__ void mark_primary(const qm_opt &o, int n, qm_reg *a, uint64_t id)
{
    if (n == 0) return;
    uint64_t hsh[QM_MAX_REGS];
    int z[QM_MAX_REGS], nz = 0;
    for (int i = 0; i < n; ++i) { a[i].sub = 0; a[i].sub_n = 0; a[i].secondary = -1; hsh[i] = hash64(id + i); }
    for (int i = 1; i < n; ++i) {
        const qm_reg x = a[i];
        const uint64_t hx = hsh[i];
        int j = i - 1;
        while (j >= 0 && !(a[j].score > x.score || (a[j].score == x.score && hsh[j] <= hx))) { a[j + 1] = a[j]; hsh[j + 1] = hsh[j]; --j; }
        a[j + 1] = x; hsh[j + 1] = hx;
    }
    int tmp = o.a + o.b;
    if (o.o_del + o.e_del > tmp) tmp = o.o_del + o.e_del;
    if (o.o_ins + o.e_ins > tmp) tmp = o.o_ins + o.e_ins;
    z[nz++] = 0;
    for (int i = 1; i < n; ++i) {
        int k;
        for (k = 0; k < nz; ++k) {
            const int j = z[k];
            const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                const int li = a[i].qe - a[i].qb, lj = a[j].qe - a[j].qb;
                const int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * o.mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i; else a[i].secondary = z[k];
    }
}

__device__ int approx_mapq(const qm_opt &o, const PairTables &T, const qm_reg &a)
{
    int sub = a.sub ? a.sub : o.min_seed_len * o.a;
    if (a.csub > sub) sub = a.csub;
    if (sub >= a.score) return 0;
    const int l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
    const double identity = 1. - (double)(l * o.a - a.score) / (o.a + o.b) / l;
    int mapq;
    if (a.score == 0) mapq = 0;
    else {
        double tmp = T.mapq_l[l < kMapqTabLen ? l : kMapqTabLen - 1];
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a.score - sub) / o.a * tmp * tmp + .499);
    }
    if (a.sub_n > 0) mapq -= T.subn[a.sub_n < kSubnTabLen ? a.sub_n : kSubnTabLen - 1];
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    return mapq;
}

__device__ __forceinline__ int raw_mapq(int diff, int a) { return (int)(6.02 * diff / a + .499); }

struct P128 { uint64_t x, y; };
__device__ __forceinline__ bool p128_gt(const P128 &a, const P128 &b) { return a.x > b.x || (a.x == b.x && a.y > b.y); }

// mem_pair.  Instead of materialising and sorting every candidate pair, keep the best two by (score, hash)
// and count the near-best ones in a second pass over the same enumeration.
__device__ int pair_up(const IndexView &V, const qm_opt &o, const PairTables &T, qm_reg *const a[2], const int n_pri[2],
                       uint64_t id, int *sub, int *n_sub, int z[2])
{
    P128 v[2 * QM_MAX_REGS];
    int nv = 0;
    const int64_t l_pac = V.l_pac;
    for (int r = 0; r < 2; ++r)
        for (int i = 0; i < n_pri[r]; ++i) {
            const qm_reg &e = a[r][i];
            const uint64_t fx = e.rb < l_pac ? (uint64_t)e.rb : (uint64_t)((l_pac << 1) - 1 - e.rb);
            v[nv].x = (uint64_t)e.rid << 32 | (fx - (uint64_t)V.off[e.rid]);
            v[nv].y = (uint64_t)e.score << 32 | (uint64_t)(i << 2) | (uint64_t)((e.rb >= l_pac) << 1) | (uint64_t)r;
            ++nv;
        }
    for (int i = 1; i < nv; ++i) { const P128 t = v[i]; int j = i - 1; while (j >= 0 && p128_gt(v[j], t)) { v[j + 1] = v[j]; --j; } v[j + 1] = t; }
    int tmp = o.a + o.b;
    if (o.o_del + o.e_del > tmp) tmp = o.o_del + o.e_del;
    if (o.o_ins + o.e_ins > tmp) tmp = o.o_ins + o.e_ins;
    P128 best = {0, 0}, second = {0, 0};
    int nu = 0, ret = 0;
    *n_sub = 0; *sub = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int y[4] = {-1, -1, -1, -1};
        for (int i = 0; i < nv; ++i) {
            for (int r = 0; r < 2; ++r) {
                const int dir = r << 1 | (int)(v[i].y >> 1 & 1);
                if (T.pes[dir].failed) continue;
                const int which = r << 1 | ((int)(v[i].y & 1) ^ 1);
                if (y[which] < 0) continue;
                for (int k = y[which]; k >= 0; --k) {
                    if ((int)(v[k].y & 3) != which) continue;
                    const int64_t dist = (int64_t)v[i].x - (int64_t)v[k].x;
                    if (dist > T.pes[dir].high) break;
                    if (dist < T.pes[dir].low) continue;
                    const double term = T.pair_term[dir][dist - T.pes[dir].low];
                    int q = (int)((double)((v[i].y >> 32) + (v[k].y >> 32)) + term + .499);
                    if (q < 0) q = 0;
                    P128 u;
                    u.y = (uint64_t)k << 32 | (uint64_t)i;
                    u.x = (uint64_t)q << 32 | (hash64(u.y ^ id << 8) & 0xffffffffU);
                    if (pass == 0) {
                        ++nu;
                        if (nu == 1 || p128_gt(u, best)) { second = best; best = u; if (nu == 1) second = u; }
                        else if (nu == 2 || p128_gt(u, second)) second = u;
                    } else if (!(u.x == best.x && u.y == best.y)) {
                        if (*sub - (int)(u.x >> 32) <= tmp) ++*n_sub;
                    }
                }
            }
            y[v[i].y & 3] = i;
        }
        if (pass == 0) {
            if (nu == 0) return 0;
            const int i = (int)(best.y >> 32), k = (int)(best.y << 32 >> 32);
            z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
            z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
            ret = (int)(best.x >> 32);
            *sub = nu > 1 ? (int)(second.x >> 32) : 0;
            if (nu == 1) break;
        }
    }
    return ret;
}

__device__ __forceinline__ int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    int w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
    if (w < abs(l1 - l2)) w = abs(l1 - l2);
    return w;
}

#define QM_NEG_INF (-0x40000000)

struct SeqPair {            // query / reference bases of one CIGAR task, reversed on the reverse strand
    const uint8_t *q;
    int lq, rlen;
    int64_t rb;
    bool rev;
    const IndexView *V;
    __device__ __forceinline__ int qb(int i) const { return rev ? q[lq - 1 - i] : q[i]; }
    __device__ __forceinline__ int tb(int i) const { return qm_ref_base(*V, rev ? rb + rlen - 1 - i : rb + i); }
};

// ksw_global2; cigar in forward order M/I/D; returns score, *n_cigar = -1 on overflow
__device__ int global_align(const qm_opt &o, const SeqPair &S, int w, int2 *eh, uint8_t *dir, int *n_cigar, uint32_t *cigar, int max_cigar)
{
    const int qlen = S.lq, tlen = S.rlen;
    const int gapo_d = o.o_del + o.e_del, gapo_i = o.o_ins + o.e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    eh[0] = make_int2(0, QM_NEG_INF);
    int col;
    for (col = 1; col <= qlen && col <= w; ++col) eh[col] = make_int2(-(o.o_ins + o.e_ins * col), QM_NEG_INF);
    for (; col <= qlen; ++col) eh[col] = make_int2(QM_NEG_INF, QM_NEG_INF);
    for (int row = 0; row < tlen; ++row) {
        int f = QM_NEG_INF;
        uint8_t *drow = dir + (size_t)row * n_col;
        const int tb = S.tb(row);
        const int lo = row > w ? row - w : 0;
        const int hi = row + w + 1 < qlen ? row + w + 1 : qlen;
        int left = lo == 0 ? -(o.o_del + o.e_del * (row + 1)) : QM_NEG_INF;
        for (col = lo; col < hi; ++col) {
            const int2 c = eh[col];
            int m = c.x, e = c.y, h, t;
            const int qb = S.qb(col);
            m += (tb > 3 || qb > 3) ? -1 : (tb == qb ? o.a : -o.b);
            uint8_t d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            t = m - gapo_d;
            e -= o.e_del;
            if (e > t) d |= 1 << 2; else e = t;
            eh[col] = make_int2(left, e);
            left = h;
            t = m - gapo_i;
            f -= o.e_ins;
            if (f > t) d |= 2 << 4; else f = t;
            drow[col - lo] = d;
        }
        eh[hi] = make_int2(left, QM_NEG_INF);
    }
    const int score = eh[qlen].x;
    int n = 0, state = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    bool overflow = false;
#define PUSH(op_, len_) do { \
        if (n > 0 && (int)(cigar[n - 1] & 0xf) == (op_)) cigar[n - 1] += (uint32_t)(len_) << 4; \
        else if (n < max_cigar) cigar[n++] = (uint32_t)(len_) << 4 | (op_); \
        else overflow = true; } while (0)
    while (i >= 0 && k >= 0) {
        const int lo = i > w ? i - w : 0;
        state = dir[(size_t)i * n_col + (k - lo)] >> (state << 1) & 3;
        if (state == 0)      { PUSH(0, 1); --i; --k; }
        else if (state == 1) { PUSH(2, 1); --i; }
        else                 { PUSH(1, 1); --k; }
    }
    if (i >= 0) PUSH(2, i + 1);
    if (k >= 0) PUSH(1, k + 1);
#undef PUSH
    for (i = 0; i < n >> 1; ++i) { const uint32_t t = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = t; }
    *n_cigar = overflow ? -1 : n;
    return score;
}

// bwa_gen_cigar2: returns score; NM via *nm
__device__ int gen_cigar(const IndexView &V, const qm_opt &o, int w_, int l_query, const uint8_t *query, int64_t rb, int64_t re,
                         uint8_t *slice, const PairScratch &PS, int *n_cigar, uint32_t *cigar, int *nm)
{
    const int64_t l_pac = V.l_pac;
    const int rlen = (int)(re - rb);
    int score = 0;
    *n_cigar = 0; *nm = -1;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
    SeqPair S;
    S.q = query; S.lq = l_query; S.rlen = rlen; S.rb = rb; S.rev = rb >= l_pac; S.V = &V;
    if (l_query == rlen && w_ == 0) {
        cigar[0] = (uint32_t)l_query << 4; *n_cigar = 1;
        for (int i = 0; i < l_query; ++i) { const int r = S.tb(i), q = S.qb(i); score += (r > 3 || q > 3) ? -1 : (r == q ? o.a : -o.b); }
    } else {
        int max_ins = (int)((double)(((l_query + 1) >> 1) * o.a - o.o_ins) / o.e_ins + 1.);
        int max_del = (int)((double)(((l_query + 1) >> 1) * o.a - o.o_del) / o.e_del + 1.);
        int max_gap = max_ins > max_del ? max_ins : max_del;
        if (max_gap < 1) max_gap = 1;
        int w = (max_gap + abs(rlen - l_query) + 1) >> 1;
        if (w > w_) w = w_;
        const int min_w = abs(rlen - l_query) + 3;
        if (w < min_w) w = min_w;
        const int n_col = l_query < 2 * w + 1 ? l_query : 2 * w + 1;
        const size_t need = (size_t)n_col * rlen;
        uint8_t *dir = slice + kEhBytes;
        if ((size_t)(l_query + 1) * sizeof(int2) > (size_t)kEhBytes) { atomicExch(PS.err, 2); *n_cigar = -1; return 0; }
        if (need > (size_t)(kSliceBytes - kEhBytes)) {
            const unsigned long long at = atomicAdd(PS.overflow_used, (unsigned long long)((need + 15) & ~(size_t)15));
            if (at + need > kOverflowBytes) { atomicExch(PS.err, 3); *n_cigar = -1; return 0; }
            dir = PS.overflow + at;
        }
        score = global_align(o, S, w, (int2 *)slice, dir, n_cigar, cigar, QM_MAX_CIGAR - 2);
    }
    if (*n_cigar > 0) {
        int x = 0, y = 0, n_mm = 0, n_gap = 0;
        for (int k = 0; k < *n_cigar; ++k) {
            const int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
            if (op == 0) { for (int i = 0; i < len; ++i) if (S.qb(x + i) != S.tb(y + i)) ++n_mm; x += len; y += len; }
            else if (op == 2) { if (k > 0 && k < *n_cigar - 1) n_gap += len; y += len; }
            else if (op == 1) { x += len; n_gap += len; }
        }
        *nm = n_mm + n_gap;
    }
    return score;
}

// mem_reg2aln
__device__ void reg_to_aln(const IndexView &V, const qm_opt &o, const PairTables &T, int l_query, const uint8_t *query,
                           const qm_reg *ar, uint8_t *slice, const PairScratch &PS, qm_aln *a)
{
    qm_aln r = {};
    if (ar == nullptr || ar->rb < 0 || ar->re < 0) { r.rid = -1; r.pos = -1; r.flag |= 0x4; *a = r; return; }
    const int qb = ar->qb, qe = ar->qe;
    const int64_t rb = ar->rb, re = ar->re;
    int nm = -1, score = 0, last_sc = -(1 << 30), n_cigar = 0;
    uint32_t cig[QM_MAX_CIGAR];
    r.mapq = ar->secondary < 0 ? (uint8_t)approx_mapq(o, T, *ar) : 0;
    if (ar->secondary >= 0) r.flag |= 0x100;
    int tmp = infer_bw(qe - qb, (int)(re - rb), ar->truesc, o.a, o.o_del, o.e_del);
    int w2 = infer_bw(qe - qb, (int)(re - rb), ar->truesc, o.a, o.o_ins, o.e_ins);
    if (tmp > w2) w2 = tmp;
    if (w2 > o.w) w2 = w2 < ar->w ? w2 : ar->w;
    int i = 0;
    do {
        if (w2 > o.w << 2) w2 = o.w << 2;
        score = gen_cigar(V, o, w2, qe - qb, query + qb, rb, re, slice, PS, &n_cigar, cig, &nm);
        if (score == last_sc || w2 == o.w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar->truesc - o.a);
    r.nm = nm;
    const bool is_rev = (rb < V.l_pac ? rb : re - 1) >= V.l_pac;
    int64_t pos = rb < V.l_pac ? rb : 2 * V.l_pac - 1 - (re - 1);
    if (n_cigar < 0) { r.rid = -1; r.pos = -1; r.flag |= 0x4; r.n_cigar = 255; *a = r; return; }
    int c0 = 0;
    if (n_cigar > 0) {
        if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; c0 = 1; --n_cigar; }
        else if ((cig[c0 + n_cigar - 1] & 0xf) == 2) --n_cigar;
    }
    int m = 0;
    const int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
    if (clip5) r.cigar[m++] = (uint32_t)clip5 << 4 | 4;
    for (i = 0; i < n_cigar; ++i) r.cigar[m++] = cig[c0 + i];
    if (clip3) r.cigar[m++] = (uint32_t)clip3 << 4 | 4;
    r.n_cigar = (uint8_t)m;
    r.rid = qm_pos2rid(V, pos);
    r.pos = (int32_t)(pos - V.off[r.rid]);
    if (is_rev) r.flag |= 0x10;
    r.score = ar->score; r.sub = ar->sub > ar->csub ? ar->sub : ar->csub;
    r.qb = qb; r.qe = qe;
    *a = r;
}

__device__ __forceinline__ int cigar_rlen(const qm_aln &a)
{
    int l = 0;
    if (a.n_cigar == 255) return 0;
    for (int k = 0; k < a.n_cigar; ++k) { const int op = a.cigar[k] & 0xf; if (op == 0 || op == 2) l += a.cigar[k] >> 4; }
    return l;
}

// flags / mate fields as bwamem.c mem_aln2sam writes them
__device__ void finish_pair(qm_aln h[2], int extra_flag)
{
    const bool mapped[2] = { h[0].rid >= 0, h[1].rid >= 0 };
    const bool rev[2] = { (h[0].flag & 0x10) != 0, (h[1].flag & 0x10) != 0 };
    const int rlen[2] = { cigar_rlen(h[0]), cigar_rlen(h[1]) };
    const int32_t pos0[2] = { h[0].pos, h[1].pos };
    const int32_t rid0[2] = { h[0].rid, h[1].rid };
    for (int i = 0; i < 2; ++i) {
        qm_aln *p = &h[i];
        const int j = !i;
        p->flag |= 0x1 | (i == 0 ? 0x40 : 0x80) | extra_flag;
        if (!mapped[j]) p->flag |= 0x8;
        if (mapped[j] && rev[j]) p->flag |= 0x20;
        if (!mapped[i] && mapped[j]) { p->rid = rid0[j]; p->pos = pos0[j]; if (rev[j]) p->flag |= 0x10; }
        if (!mapped[i] && !mapped[j]) { p->mate_rid = -1; p->mate_pos = -1; p->tlen = 0; continue; }
        if (mapped[j]) { p->mate_rid = rid0[j]; p->mate_pos = pos0[j]; }
        else { p->mate_rid = rid0[i]; p->mate_pos = pos0[i]; if (rev[i]) p->flag |= 0x20; }
        p->tlen = 0;
        if (mapped[0] && mapped[1] && rid0[0] == rid0[1]) {
            const int64_t p0 = pos0[i] + (rev[i] ? rlen[i] - 1 : 0);
            const int64_t p1 = pos0[j] + (rev[j] ? rlen[j] - 1 : 0);
            p->tlen = (int32_t)(-(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
        }
    }
}

__global__ void __launch_bounds__(128)
pair_kernel(IndexView V, qm_opt o, PairTables T, PairScratch PS, const uint8_t *__restrict__ codes, int stride,
            const int32_t *__restrict__ lens, int64_t n_pairs, int64_t pair_id0, qm_reg *__restrict__ regs,
            int32_t *__restrict__ n_regs, qm_aln *__restrict__ alns)
{
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    uint8_t *slice = PS.slices + tid * kSliceBytes;
    for (int64_t pi = tid; pi < n_pairs; pi += nthreads) {
        qm_reg *a[2] = { regs + (2 * pi) * QM_MAX_REGS, regs + (2 * pi + 1) * QM_MAX_REGS };
        const int n[2] = { n_regs[2 * pi], n_regs[2 * pi + 1] };
        int n_pri[2], z[2] = {0, 0};
        const uint8_t *seq[2] = { codes + (2 * pi) * stride, codes + (2 * pi + 1) * stride };
        const int l_seq[2] = { lens[2 * pi], lens[2 * pi + 1] };
        const uint64_t id = (uint64_t)(pair_id0 + pi);
        qm_aln h[2];
        int extra_flag = 0, o_sc = 0, subo = 0, n_sub = 0;
        bool paired_done = false;
        mark_primary(o, n[0], a[0], id << 1 | 0);
        mark_primary(o, n[1], a[1], id << 1 | 1);
        n_pri[0] = n[0]; n_pri[1] = n[1];
        if (n_pri[0] && n_pri[1] && (o_sc = pair_up(V, o, T, a, n_pri, id, &subo, &n_sub, z)) > 0) {
            bool is_multi[2];
            for (int i = 0; i < 2; ++i) {
                int j;
                for (j = 1; j < n_pri[i]; ++j) if (a[i][j].secondary < 0 && a[i][j].score >= o.T) break;
                is_multi[i] = j < n_pri[i];
            }
            if (!(is_multi[0] || is_multi[1])) {
                int q_se[2];
                const int score_un = a[0][0].score + a[1][0].score - o.pen_unpaired;
                if (score_un > subo) subo = score_un;
                int q_pe = raw_mapq(o_sc - subo, o.a);
                if (n_sub > 0) q_pe -= T.subn[n_sub < kSubnTabLen ? n_sub : kSubnTabLen - 1];
                if (q_pe < 0) q_pe = 0;
                if (q_pe > 60) q_pe = 60;
                if (o_sc > score_un) {
                    qm_reg *c[2] = { &a[0][z[0]], &a[1][z[1]] };
                    for (int i = 0; i < 2; ++i) {
                        if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
                        q_se[i] = approx_mapq(o, T, *c[i]);
                    }
                    for (int i = 0; i < 2; ++i) {
                        q_se[i] = q_se[i] > q_pe ? q_se[i] : q_pe < q_se[i] + 40 ? q_pe : q_se[i] + 40;
                        const int cap = raw_mapq(c[i]->score - c[i]->csub, o.a);
                        if (q_se[i] > cap) q_se[i] = cap;
                    }
                    extra_flag |= 2;
                } else {
                    z[0] = z[1] = 0;
                    q_se[0] = approx_mapq(o, T, a[0][0]);
                    q_se[1] = approx_mapq(o, T, a[1][0]);
                }
                for (int i = 0; i < 2; ++i) {
                    reg_to_aln(V, o, T, l_seq[i], seq[i], &a[i][z[i]], slice, PS, &h[i]);
                    h[i].mapq = (uint8_t)q_se[i];
                    h[i].flag &= ~0x100;
                }
                paired_done = true;
            }
        }
        if (!paired_done) {
            for (int i = 0; i < 2; ++i) {
                if (n[i] && a[i][0].score >= o.T) reg_to_aln(V, o, T, l_seq[i], seq[i], &a[i][0], slice, PS, &h[i]);
                else reg_to_aln(V, o, T, l_seq[i], seq[i], nullptr, slice, PS, &h[i]);
            }
            if (h[0].rid == h[1].rid && h[0].rid >= 0) {
                int64_t dist;
                const int d = infer_dir(V.l_pac, a[0][0].rb, a[1][0].rb, &dist);
                if (!T.pes[d].failed && dist >= T.pes[d].low && dist <= T.pes[d].high) extra_flag |= 2;
            }
        }
        finish_pair(h, extra_flag);
        alns[2 * pi] = h[0]; alns[2 * pi + 1] = h[1];
    }
}

}  // namespace

extern "C" {

int qm_pestat_sync(qm_ctx *ctx, const qm_index *idx, const qm_opt *opt, const qm_reg *d_regs, const int32_t *d_n_regs,
                   int64_t n_pairs, qm_pestat pes[4], void *stream)
{
    if (!ctx || !idx || !opt || !pes || n_pairs < 0 || (n_pairs > 0 && (!d_regs || !d_n_regs))) return QM_EINVAL;
    if (opt->max_ins < 1 || opt->max_ins > (1 << 20)) return qm_fail(ctx, QM_ELIMIT, "qm_pestat_sync: max_ins must be in [1, 2^20]");
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // device: one histogram of insert sizes per orientation; host: the percentile / moment arithmetic of
    // mem_pestat replayed over the histogram in ascending order (same operation order as over the sorted list)
    const size_t bins = (size_t)opt->max_ins + 1;
    std::vector<unsigned> h(4 * bins, 0u);
    if (n_pairs > 0) {
        void *p = nullptr;
        int rc = qm_scratch_reserve(ctx, 4, 4 * bins * sizeof(unsigned), &p);
        if (rc) return rc;
        QM_CUDA(ctx, cudaMemsetAsync(p, 0, 4 * bins * sizeof(unsigned), st));
        const int tpb = 128;
        const int sp = qm_prof_begin(ctx, QM_ST_PAIR, st);
        pestat_hist_kernel<<<(unsigned)((n_pairs + tpb - 1) / tpb), tpb, 0, st>>>(idx->v, *opt, d_regs, d_n_regs, n_pairs, (unsigned *)p);
        qm_prof_end(ctx, QM_ST_PAIR, sp, st, 1);
        QM_CUDA(ctx, cudaMemcpyAsync(h.data(), p, 4 * bins * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        QM_CUDA(ctx, cudaStreamSynchronize(st));
    }
    int64_t cnt[4], max = 0;
    for (int d = 0; d < 4; ++d) {
        qm_pestat *r = &pes[d];
        r->low = r->high = r->failed = r->pad = 0; r->avg = r->std = 0;
        const unsigned *q = h.data() + (size_t)d * bins;
        int64_t n = 0;
        for (size_t v = 0; v < bins; ++v) n += q[v];
        cnt[d] = n;
        if (n > max) max = n;
        if (n < 10) { r->failed = 1; continue; }
        auto nth = [&](int64_t k) {       // value at rank k of the ascending list
            int64_t c = 0;
            for (size_t v = 0; v < bins; ++v) { c += q[v]; if (c > k) return (int)v; }
            return (int)bins - 1;
        };
        const int p25 = nth((int)(.25 * n + .499)), p75 = nth((int)(.75 * n + .499));
        r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
        if (r->low < 1) r->low = 1;
        r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
        int64_t x = 0;
        for (int64_t v = r->low; v <= r->high && v < (int64_t)bins; ++v) for (unsigned c = 0; c < q[v]; ++c) { r->avg += (double)v; ++x; }
        r->avg /= x;
        for (int64_t v = r->low; v <= r->high && v < (int64_t)bins; ++v) {
            const double dlt = ((double)v - r->avg) * ((double)v - r->avg);
            for (unsigned c = 0; c < q[v]; ++c) r->std += dlt;
        }
        r->std = sqrt(r->std / x);
        r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
        r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
        if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
        if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
        if (r->low < 1) r->low = 1;
    }
    for (int d = 0; d < 4; ++d) if (!pes[d].failed && cnt[d] < max * 0.05) pes[d].failed = 1;
    return QM_OK;
}

int qm_pair_finish(qm_ctx *ctx, const qm_index *idx, const qm_opt *opt, const uint8_t *d_codes, int32_t stride,
                   const int32_t *d_lens, int64_t n_pairs, int64_t pair_id0, qm_reg *d_regs, int32_t *d_n_regs,
                   const qm_pestat pes[4], qm_aln *d_alns, void *stream)
{
    if (!ctx || !idx || !opt || !pes || n_pairs < 0 || (n_pairs > 0 && (!d_codes || !d_lens || !d_regs || !d_n_regs || !d_alns)))
        return QM_EINVAL;
    if (n_pairs == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    // host-evaluated tables (libm): MAPQ length factor, sub_n penalty, pairing term per insert size
    std::vector<double> tab(kMapqTabLen);
    for (int l = 0; l < kMapqTabLen; ++l)
        tab[l] = l < opt->mapq_coef_len ? 1. : log((double)opt->mapq_coef_len) / log((double)l);
    std::vector<int> subn(kSubnTabLen);
    for (int n = 0; n < kSubnTabLen; ++n) subn[n] = (int)(4.343 * log((double)(n + 1)) + .499);
    size_t term_off[4], n_term = 0;
    for (int d = 0; d < 4; ++d) {
        term_off[d] = n_term;
        if (!pes[d].failed && pes[d].high >= pes[d].low) n_term += (size_t)(pes[d].high - pes[d].low + 1);
    }
    std::vector<double> term(n_term + 1);
    for (int d = 0; d < 4; ++d) {
        if (pes[d].failed || pes[d].high < pes[d].low) continue;
        for (int64_t dist = pes[d].low; dist <= pes[d].high; ++dist) {
            const double ns = (dist - pes[d].avg) / pes[d].std;
            term[term_off[d] + (size_t)(dist - pes[d].low)] = .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * opt->a;
        }
    }
    const int threads_total = ctx->sm_count * 256;
    const size_t o_tab = 0, o_subn = o_tab + kMapqTabLen * 8, o_term = o_subn + kSubnTabLen * 4;
    const size_t o_misc = (o_term + term.size() * 8 + 255) & ~(size_t)255;
    const size_t o_slices = o_misc + 256;
    const size_t o_over = o_slices + (size_t)threads_total * kSliceBytes;
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 5, o_over + kOverflowBytes, &p);
    if (rc) return rc;
    char *b = (char *)p;
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_tab, tab.data(), kMapqTabLen * 8, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_subn, subn.data(), kSubnTabLen * 4, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_term, term.data(), term.size() * 8, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemsetAsync(b + o_misc, 0, 256, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));      // the host vectors above go out of scope
    PairTables T;
    T.mapq_l = (const double *)(b + o_tab); T.subn = (const int *)(b + o_subn);
    for (int d = 0; d < 4; ++d) { T.pair_term[d] = (const double *)(b + o_term) + term_off[d]; T.pes[d] = pes[d]; }
    PairScratch PS;
    PS.slices = (uint8_t *)(b + o_slices); PS.overflow = (uint8_t *)(b + o_over);
    PS.overflow_used = (unsigned long long *)(b + o_misc); PS.err = (int *)(b + o_misc + 16);
    int64_t blocks = (n_pairs + 127) / 128;
    if (blocks > threads_total / 128) blocks = threads_total / 128;
    const int sp = qm_prof_begin(ctx, QM_ST_PAIR, st);
    pair_kernel<<<(unsigned)blocks, 128, 0, st>>>(idx->v, *opt, T, PS, d_codes, stride, d_lens, n_pairs, pair_id0, d_regs, d_n_regs, d_alns);
    qm_prof_end(ctx, QM_ST_PAIR, sp, st, 1);
    QM_CUDA(ctx, cudaGetLastError());
    int h_err = 0;
    QM_CUDA(ctx, cudaMemcpyAsync(&h_err, PS.err, 4, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_err) return qm_fail(ctx, QM_ELIMIT, "qm_pair_finish: traceback scratch exhausted (code %d)", h_err);
    return QM_OK;
}

}  // extern "C"
