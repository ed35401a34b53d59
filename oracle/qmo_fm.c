/* qmo_fm.c -- CPU restatement of bwa 0.7.17's FM-index seeding (TEST INFRASTRUCTURE, like the rest of oracle/):
 * index construction as `bwa index` does it (rules/index.smk:13: bwt_pac2bwt -> is_bwt, bwt_bwtupdate_core, bwt_cal_sa),
 * the bidirectional extension (bwt.c: bwt_occ4 / bwt_2occ4 / bwt_extend), SMEM search (bwt_smem1a), the three seeding
 * rounds of bwamem.c mem_collect_intv (SMEMs, re-seeding of long unique SMEMs, the LAST-like third round
 * bwt_seed_strategy1) and the suffix-array look-up (bwt_sa / bwt_invPsi) that turns intervals into seeds the way
 * mem_chain walks them.  The upstream sources are not vendored: written from the published
 * algorithm (Li 2012, "Exploring single-sample SNP and INDEL calling with whole-genome de novo assembly", and the
 * bwa-mem manuscript).
 *
 * PINNED where the reference allows it: the index this file builds from the packed genome equals, word for word, bwa's
 * own index files shipped in the reference (ref/X.bwt incl. the interleaved occurrence checkpoints, ref/X.sa) -- digests
 * in tests/golden/fm_digests.json, made by tests/golden/make_fm_digests.py from /root/reference/ref.  The search
 * functions then run on the reference's own bytes (qmo_fm_load).  What stays unpinned is the search logic itself: no
 * bwa binary exists here to compare seeds with. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "qmo_priv.h"

#define OCC_INTV 128

typedef struct {
    int64_t primary, L2[5], seq_len;       /* seq_len = 2 l_pac: forward strand then its reverse complement */
    int64_t n_words;                       /* uint32 words of the interleaved array */
    uint32_t *bwt;                         /* per 128 bases: 4 x int64 occurrence counts (8 words), then 8 words of 16 bases */
    int sa_intv;
    int64_t n_sa;
    int64_t *sa;                           /* sa[k / sa_intv] for k % sa_intv == 0, k = row of the full (n + 1)-row matrix; sa[0] = -1 */
} qmo_fm_t;

/* ---- suffix array of T[0..n) by prefix doubling (ranks compared pairwise, qsort per round) ---- */
static const int64_t *g_rank;
static int64_t g_k, g_n;
static int sa_cmp(const void *a, const void *b)
{
    const int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    if (g_rank[x] != g_rank[y]) return g_rank[x] < g_rank[y] ? -1 : 1;
    {
        const int64_t rx = x + g_k < g_n ? g_rank[x + g_k] : -1, ry = y + g_k < g_n ? g_rank[y + g_k] : -1;
        return rx < ry ? -1 : rx > ry;
    }
}
/* doubling with re-sorting restricted to the groups that are still tied */
static int64_t *suffix_array(const uint8_t *T, int64_t n)
{
    int64_t *sa = (int64_t *)malloc(8 * (size_t)(n + 1)), *rank = (int64_t *)malloc(8 * (size_t)(n + 1)), *tmp = (int64_t *)malloc(8 * (size_t)(n + 1));
    int64_t i, k;
    /* initial ranks: the first 8 bases as a number (an absent base sorts below A) */
    for (i = 0; i < n; ++i) {
        int64_t v = 0; int j;
        for (j = 0; j < 8; ++j) v = v * 5 + (i + j < n ? T[i + j] + 1 : 0);
        rank[i] = v; sa[i] = i;
    }
    g_rank = rank; g_n = n; g_k = 0;
    /* counting is overkill here: one qsort on the 8-base keys */
    {
        g_k = n;                                   /* second key off */
        qsort(sa, (size_t)n, 8, sa_cmp);
    }
    /* dense ranks */
    tmp[sa[0]] = 0;
    for (i = 1; i < n; ++i) tmp[sa[i]] = tmp[sa[i - 1]] + (rank[sa[i]] != rank[sa[i - 1]]);
    memcpy(rank, tmp, 8 * (size_t)n);
    for (k = 8; k < n; k <<= 1) {
        int64_t b = 0;
        int any = 0;
        g_k = k;
        while (b < n) {
            int64_t e = b + 1;
            while (e < n && rank[sa[e]] == rank[sa[b]]) ++e;
            if (e - b > 1) { qsort(sa + b, (size_t)(e - b), 8, sa_cmp); any = 1; }
            b = e;
        }
        if (!any) break;
        tmp[sa[0]] = 0;
        for (i = 1; i < n; ++i) {
            const int64_t x = sa[i - 1], y = sa[i];
            const int64_t rx = x + k < n ? rank[x + k] : -1, ry = y + k < n ? rank[y + k] : -1;
            tmp[y] = tmp[x] + (rank[x] != rank[y] || rx != ry);
        }
        memcpy(rank, tmp, 8 * (size_t)n);
        if (rank[sa[n - 1]] == n - 1) break;
    }
    free(rank); free(tmp);
    return sa;
}

static inline uint32_t fm_word(const qmo_fm_t *F, int64_t k) { return F->bwt[(k >> 7 << 4) + 8 + ((k & 0x7f) >> 4)]; }
static inline int fm_base(const qmo_fm_t *F, int64_t k) { return (int)(fm_word(F, k) >> ((~k & 0xf) << 1) & 3); }

/* `bwa index`: BWT of forward + reverse complement with the sentinel row dropped (is_bwt), occurrence checkpoints every
 * 128 bases interleaved with the 2-bit words (bwt_bwtupdate_core), suffix-array samples every sa_intv rows (bwt_cal_sa) */
qmo_fm_t *qmo_fm_build(const uint8_t *fwd, int64_t l_pac, int sa_intv)
{
    const int64_t n = 2 * l_pac;
    uint8_t *T = (uint8_t *)malloc((size_t)n);
    int64_t i, *sa, primary = 0, c4[4] = {0, 0, 0, 0};
    qmo_fm_t *F = (qmo_fm_t *)calloc(1, sizeof(qmo_fm_t));
    uint8_t *B;
    for (i = 0; i < l_pac; ++i) { T[i] = fwd[i]; T[n - 1 - i] = (uint8_t)(3 - fwd[i]); }
    sa = suffix_array(T, n);                       /* rows 1..n of the matrix; row 0 is the sentinel suffix */
    B = (uint8_t *)malloc((size_t)n);
    {
        int64_t w = 0, r;
        for (r = 0; r <= n; ++r) {                 /* full matrix: row 0 is the sentinel suffix, row r > 0 suffix sa[r - 1] */
            const int64_t sfx = r == 0 ? n : sa[r - 1];
            if (sfx == 0) { primary = r; continue; }
            B[w++] = T[sfx - 1];
        }
    }
    F->primary = primary; F->seq_len = n; F->sa_intv = sa_intv;
    for (i = 0; i < n; ++i) ++c4[B[i]];
    F->L2[0] = 0;
    for (i = 0; i < 4; ++i) F->L2[i + 1] = F->L2[i] + c4[i];
    {
        const int64_t n_occ = (n + OCC_INTV - 1) / OCC_INTV + 1, plain = (n + 15) >> 4;
        int64_t k = 0, cnt[4] = {0, 0, 0, 0};
        F->n_words = plain + n_occ * 8;
        F->bwt = (uint32_t *)calloc((size_t)F->n_words, 4);
        for (i = 0; i < n; ++i) {
            if (i % OCC_INTV == 0) { memcpy(F->bwt + k, cnt, 32); k += 8; }
            if (i % 16 == 0) ++k;
            F->bwt[k - 1] |= (uint32_t)B[i] << ((~i & 0xf) << 1);
            ++cnt[B[i]];
        }
        memcpy(F->bwt + k, cnt, 32);               /* the last checkpoint */
        k += 8;
        if (k != F->n_words) { /* n a multiple of 128: bwa writes the trailing checkpoint all the same */ F->n_words = k; }
    }
    /* samples: row r of the full matrix holds suffix (r == 0 ? n : sa[r - 1]) */
    F->n_sa = (n + sa_intv) / sa_intv;
    F->sa = (int64_t *)calloc((size_t)F->n_sa, 8);
    for (i = 0; i <= n; i += sa_intv) F->sa[i / sa_intv] = i == 0 ? -1 : sa[i - 1];
    free(sa); free(B); free(T);
    return F;
}

/* bwa's own files: .bwt = primary, L2[1..4], interleaved words; .sa = primary, 4 skipped words, sa_intv, seq_len, samples 1.. */
qmo_fm_t *qmo_fm_load(const uint8_t *bwt_bytes, int64_t bwt_len, const uint8_t *sa_bytes, int64_t sa_len)
{
    qmo_fm_t *F = (qmo_fm_t *)calloc(1, sizeof(qmo_fm_t));
    int64_t hdr[5], sh[7];
    if (bwt_len < 40 || sa_len < 56) { free(F); return 0; }
    memcpy(hdr, bwt_bytes, 40);
    F->primary = hdr[0]; F->L2[0] = 0; memcpy(F->L2 + 1, hdr + 1, 32);
    F->seq_len = F->L2[4];
    F->n_words = (bwt_len - 40) / 4;
    F->bwt = (uint32_t *)malloc((size_t)F->n_words * 4);
    memcpy(F->bwt, bwt_bytes + 40, (size_t)F->n_words * 4);
    memcpy(sh, sa_bytes, 56);
    if (sh[0] != F->primary || sh[6] != F->seq_len) { free(F->bwt); free(F); return 0; }
    F->sa_intv = (int)sh[5];
    F->n_sa = (F->seq_len + F->sa_intv) / F->sa_intv;
    if (sa_len < 56 + 8 * (F->n_sa - 1)) { free(F->bwt); free(F); return 0; }
    F->sa = (int64_t *)malloc((size_t)F->n_sa * 8);
    F->sa[0] = -1;
    memcpy(F->sa + 1, sa_bytes + 56, (size_t)(F->n_sa - 1) * 8);
    return F;
}
void qmo_fm_free(qmo_fm_t *F) { if (F) { free(F->bwt); free(F->sa); free(F); } }
/* serialised exactly as bwa writes the two files (for the digests) */
int64_t qmo_fm_bwt_bytes(const qmo_fm_t *F, uint8_t *out)
{
    if (out) { int64_t hdr[5] = {F->primary, F->L2[1], F->L2[2], F->L2[3], F->L2[4]}; memcpy(out, hdr, 40); memcpy(out + 40, F->bwt, (size_t)F->n_words * 4); }
    return 40 + F->n_words * 4;
}
int64_t qmo_fm_sa_bytes(const qmo_fm_t *F, uint8_t *out)
{
    if (out) {
        int64_t sh[7] = {F->primary, F->L2[1], F->L2[2], F->L2[3], F->L2[4], F->sa_intv, F->seq_len};
        memcpy(out, sh, 56); memcpy(out + 56, F->sa + 1, (size_t)(F->n_sa - 1) * 8);
    }
    return 56 + (F->n_sa - 1) * 8;
}

/* ---- occurrence counts (bwt.c bwt_occ4: the counts of all four bases in rows [0, k], k in matrix rows) ---- */
static void fm_occ4(const qmo_fm_t *F, int64_t k, int64_t cnt[4])
{
    int64_t j, end;
    const uint32_t *p;
    if (k == -1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    k -= (k >= F->primary);                        /* the sentinel row is not stored */
    p = F->bwt + (k >> 7 << 4);
    memcpy(cnt, p, 32);
    end = k + 1;                                   /* count stored positions [block start, k] */
    for (j = k >> 7 << 7; j < end; ++j) ++cnt[fm_base(F, j)];
}
typedef struct { int64_t x0, x1, x2; int64_t info; } fm_intv;

static void fm_set_intv(const qmo_fm_t *F, int c, fm_intv *ik)
{
    ik->x0 = F->L2[c] + 1; ik->x2 = F->L2[c + 1] - F->L2[c]; ik->x1 = F->L2[3 - c] + 1; ik->info = 0;
}
/* bwt.c bwt_extend: the four one-base extensions of a bi-interval, backward (is_back) or forward */
static void fm_extend(const qmo_fm_t *F, const fm_intv *ik, fm_intv ok[4], int is_back)
{
    int64_t tk[4], tl[4];
    const int64_t a = is_back ? ik->x0 : ik->x1, b = is_back ? ik->x1 : ik->x0;
    int i;
    fm_occ4(F, a - 1, tk);
    fm_occ4(F, a - 1 + ik->x2, tl);
    for (i = 0; i < 4; ++i) {
        const int64_t na = F->L2[i] + 1 + tk[i];
        ok[i].x2 = tl[i] - tk[i];
        if (is_back) ok[i].x0 = na; else ok[i].x1 = na;
    }
    {
        int64_t v = b + (a <= F->primary && a + ik->x2 - 1 >= F->primary);
        int i2;
        for (i2 = 3; i2 >= 0; --i2) { if (is_back) ok[i2].x1 = v; else ok[i2].x0 = v; v += ok[i2].x2; }
    }
}

/* bwt.c bwt_smem1a with max_intv = 0 (what bwt_smem1 passes): all SMEMs through position x, leftmost first */
#define QMO_FM_MAXV 48            /* interval sizes one forward search can pass through (limit shared with the device code) */
#define QMO_FM_MAXIV 96           /* intervals collected per read over the three rounds (shared limit) */
static int fm_smem1(const qmo_fm_t *F, int len, const uint8_t *q, int x, int64_t min_intv, fm_intv *mem, int *n_mem, int max_mem)
{
    fm_intv ik, ok[4], va[QMO_FM_MAXV], vb[QMO_FM_MAXV], *prev = va, *curr = vb, *swap;
    int i, j, c, ret, n_prev = 0, n_curr = 0;
    *n_mem = 0;
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    fm_set_intv(F, q[x], &ik);
    ik.info = x + 1;
    for (i = x + 1; i < len; ++i) {                /* forward: every interval size the match goes through */
        if (q[i] < 4) {
            c = 3 - q[i];
            fm_extend(F, &ik, ok, 0);
            if (ok[c].x2 != ik.x2) {
                if (n_curr < QMO_FM_MAXV) curr[n_curr++] = ik;
                if (ok[c].x2 < min_intv) break;
            }
            ik = ok[c]; ik.info = i + 1;
        } else { if (n_curr < QMO_FM_MAXV) curr[n_curr++] = ik; break; }
    }
    if (i == len && n_curr < QMO_FM_MAXV) curr[n_curr++] = ik;
    for (i = 0; i < n_curr >> 1; ++i) { fm_intv t = curr[i]; curr[i] = curr[n_curr - 1 - i]; curr[n_curr - 1 - i] = t; }   /* longest first */
    ret = (int)curr[0].info;
    swap = curr; curr = prev; prev = swap; n_prev = n_curr;
    for (i = x - 1; i >= -1; --i) {                /* backward: extend all of them, keep what cannot go on */
        c = i < 0 ? -1 : q[i] < 4 ? q[i] : -1;
        for (j = 0, n_curr = 0; j < n_prev; ++j) {
            fm_intv *p = &prev[j];
            if (c >= 0) fm_extend(F, p, ok, 1);
            if (c < 0 || ok[c].x2 < min_intv) {
                if (n_curr == 0) {
                    if (*n_mem == 0 || i + 1 < (int)(mem[*n_mem - 1].info >> 32)) {
                        ik = *p; ik.info |= (int64_t)(i + 1) << 32;
                        if (*n_mem < max_mem) mem[(*n_mem)++] = ik;
                    }
                }
            } else if (n_curr == 0 || ok[c].x2 != curr[n_curr - 1].x2) {
                ok[c].info = p->info;
                curr[n_curr++] = ok[c];
            }
        }
        if (n_curr == 0) break;
        swap = curr; curr = prev; prev = swap; n_prev = n_curr;
    }
    for (i = 0; i < *n_mem >> 1; ++i) { fm_intv t = mem[i]; mem[i] = mem[*n_mem - 1 - i]; mem[*n_mem - 1 - i] = t; }       /* by start */
    return ret;
}
/* bwt.c bwt_seed_strategy1 (third round): the shortest match from x of at least min_len bases with fewer than max_intv hits */
static int fm_seed_strategy1(const qmo_fm_t *F, int len, const uint8_t *q, int x, int min_len, int max_intv, fm_intv *mem)
{
    fm_intv ik, ok[4];
    int i, c;
    memset(mem, 0, sizeof *mem);
    if (q[x] > 3) return x + 1;
    fm_set_intv(F, q[x], &ik);
    for (i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            c = 3 - q[i];
            fm_extend(F, &ik, ok, 0);
            if (ok[c].x2 < max_intv && i - x >= min_len) { *mem = ok[c]; mem->info = (int64_t)x << 32 | (i + 1); return i + 1; }
            ik = ok[c];
        } else return i + 1;
    }
    return len;
}
static int intv_cmp(const void *a, const void *b)
{
    const uint64_t x = (uint64_t)((const fm_intv *)a)->info, y = (uint64_t)((const fm_intv *)b)->info;
    return x < y ? -1 : x > y;
}
/* bwt.c bwt_sa / bwt_invPsi: the text position of matrix row k */
static int64_t fm_sa(const qmo_fm_t *F, int64_t k)
{
    int64_t sa = 0;
    const int64_t mask = F->sa_intv - 1;
    while (k & mask) {
        int64_t cnt[4], x;
        ++sa;
        if (k == F->primary) { k = 0; continue; }
        x = fm_base(F, k - (k > F->primary));
        fm_occ4(F, k, cnt);
        k = F->L2[x] + cnt[x];
    }
    return sa + F->sa[k / F->sa_intv];
}

/* bwamem.c mem_collect_intv + the seed loop of mem_chain: intervals of the three rounds sorted by (start, end) of the match,
 * each turned into at most max_occ seeds (every step-th occurrence), seeds that bridge two contigs or the strand boundary
 * dropped.  seeds: {rbeg (doubled coordinates), qbeg, len} in bwa's order; returns their number (<= max_seeds are written). */
int qmo_fm_seeds(const void *Fv, const qmo_ref_t *R, const qmo_opt_t *o, int len, const uint8_t *q, int max_mem_intv,
                 int64_t *seeds /* 3 per seed */, int max_seeds)
{
    const qmo_fm_t *F = (const qmo_fm_t *)Fv;
    fm_intv mem[QMO_FM_MAXIV], m1[QMO_FM_MAXIV];
    int n = 0, n1, x = 0, i, k, old_n, ns = 0;
    const int split_len = (int)(o->min_seed_len * 1.5 + .499), split_width = 10;
    while (x < len) {
        if (q[x] < 4) {
            x = fm_smem1(F, len, q, x, 1, m1, &n1, QMO_FM_MAXIV - n);
            for (i = 0; i < n1; ++i)
                if ((int)(uint32_t)m1[i].info - (int)(m1[i].info >> 32) >= o->min_seed_len && n < QMO_FM_MAXIV) mem[n++] = m1[i];
        } else ++x;
    }
    old_n = n;
    for (k = 0; k < old_n; ++k) {
        const int start = (int)(mem[k].info >> 32), end = (int)(uint32_t)mem[k].info;
        if (end - start < split_len || mem[k].x2 > split_width) continue;
        fm_smem1(F, len, q, (start + end) >> 1, mem[k].x2 + 1, m1, &n1, QMO_FM_MAXIV - n);
        for (i = 0; i < n1; ++i)
            if ((int)(uint32_t)m1[i].info - (int)(m1[i].info >> 32) >= o->min_seed_len && n < QMO_FM_MAXIV) mem[n++] = m1[i];
    }
    if (max_mem_intv > 0) {
        x = 0;
        while (x < len) {
            if (q[x] < 4) {
                fm_intv m;
                x = fm_seed_strategy1(F, len, q, x, o->min_seed_len, max_mem_intv, &m);
                if (m.x2 > 0 && n < QMO_FM_MAXIV) mem[n++] = m;
            } else ++x;
        }
    }
    qsort(mem, (size_t)n, sizeof(fm_intv), intv_cmp);
    for (i = 0; i < n; ++i) {
        const fm_intv *p = &mem[i];
        const int slen = (int)(uint32_t)p->info - (int)(p->info >> 32), qbeg = (int)(p->info >> 32);
        const int64_t step = p->x2 > o->max_occ ? p->x2 / o->max_occ : 1;
        int64_t kk;
        int count;
        for (kk = 0, count = 0; kk < p->x2 && count < o->max_occ; kk += step, ++count) {
            const int64_t rbeg = fm_sa(F, p->x0 + kk), rend = rbeg + slen;
            /* bns_intv2rid: both ends on one strand and inside one contig */
            int c, ok = 0;
            if (!(rbeg < R->l_pac && rend > R->l_pac)) {
                const int64_t fb = rbeg >= R->l_pac ? 2 * R->l_pac - rend : rbeg, fe = fb + slen;
                for (c = 0; c < R->n_contigs; ++c) if (fb >= R->off[c] && fe <= R->off[c] + R->len[c]) ok = 1;
            }
            if (!ok) continue;
            if (ns < max_seeds) { seeds[3 * ns] = rbeg; seeds[3 * ns + 1] = qbeg; seeds[3 * ns + 2] = slen; }
            ++ns;
        }
    }
    return ns;
}
