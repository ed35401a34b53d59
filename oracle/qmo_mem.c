/*
 * qmo_mem.c -- ORACLE (test infrastructure): CPU restatement of the bwa-mem glue around the DP kernels
 * for the read-level hot path (reference call site rules/bwa.smk:15 `bwa mem -k 31`; upstream
 * bwamem.c / bwamem_pair.c / bwa.c are not vendored -- SURVEY.md Appendix A.1-A.6 is the spec).
 * PARITY UNPINNED -- see qmo.h.
 *
 * Seeding differs from bwa BY DESIGN (BASELINE.json north_star: "k-mer hash index"): seeds are all
 * maximal exact matches of length >= min_seed_len between the read and either strand of the
 * reference, found through a k-mer index with k = min_seed_len (SURVEY.md A.2 first bullet).  The
 * oracle looks k-mers up by binary search in a sorted array; the product uses a hash table.
 * Documented simplifications shared with the product (DESIGN.md "deviations"): no re-seeding, no
 * mem_patch_reg, frac_rep = 0, csub = 0 for extended regions, primary alignment only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "qmo_priv.h"


typedef struct { uint64_t key; uint32_t pos; } kmpair_t;
static int kmpair_cmp(const void *a, const void *b)
{
    const kmpair_t *x = (const kmpair_t *)a, *y = (const kmpair_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : x->pos > y->pos;
}

qmo_ref_t *qmo_ref_create(const uint8_t *codes, int n_contigs, const int64_t *lens, int k)
{
    qmo_ref_t *r = (qmo_ref_t *)calloc(1, sizeof(*r));
    int c;
    int64_t p, n = 0, total = 0;
    r->n_contigs = n_contigs; r->k = k;
    r->off = (int64_t *)calloc(n_contigs, 8); r->len = (int64_t *)calloc(n_contigs, 8);
    for (c = 0; c < n_contigs; ++c) { r->off[c] = total; r->len[c] = lens[c]; total += lens[c]; }
    r->l_pac = total;
    r->fwd = (uint8_t *)malloc(total);
    memcpy(r->fwd, codes, total);
    kmpair_t *tmp = (kmpair_t *)malloc(sizeof(kmpair_t) * (size_t)(total + 1));
    const uint64_t mask = k < 32 ? ((1ULL << (2 * k)) - 1) : ~0ULL;
    for (c = 0; c < n_contigs; ++c) {
        uint64_t key = 0;
        for (p = 0; p < lens[c]; ++p) {
            key = ((key << 2) | codes[r->off[c] + p]) & mask;
            if (p >= k - 1) { tmp[n].key = key; tmp[n].pos = (uint32_t)(r->off[c] + p - (k - 1)); ++n; }
        }
    }
    qsort(tmp, n, sizeof(kmpair_t), kmpair_cmp);
    r->n_km = n;
    r->km_key = (uint64_t *)malloc(8 * (size_t)(n + 1));
    r->km_pos = (uint32_t *)malloc(4 * (size_t)(n + 1));
    for (p = 0; p < n; ++p) { r->km_key[p] = tmp[p].key; r->km_pos[p] = tmp[p].pos; }
    free(tmp);
    return r;
}

void qmo_ref_destroy(qmo_ref_t *r)
{
    if (!r) return;
    free(r->off); free(r->len); free(r->fwd); free(r->km_key); free(r->km_pos); free(r);
}
int64_t qmo_ref_lpac(const qmo_ref_t *r) { return r->l_pac; }

static inline int ref_base(const qmo_ref_t *R, int64_t x)
{   /* doubled coordinates: [l_pac, 2 l_pac) is the reverse complement */
    return x < R->l_pac ? R->fwd[x] : 3 - R->fwd[2 * R->l_pac - 1 - x];
}
static inline int64_t depos(const qmo_ref_t *R, int64_t x, int *is_rev)
{
    *is_rev = x >= R->l_pac;
    return *is_rev ? 2 * R->l_pac - 1 - x : x;
}
static int pos2rid(const qmo_ref_t *R, int64_t fpos)
{
    int c;
    for (c = 0; c < R->n_contigs; ++c)
        if (fpos >= R->off[c] && fpos < R->off[c] + R->len[c]) return c;
    return -1;
}

/* first index with key >= x */
static int64_t km_lower(const qmo_ref_t *R, uint64_t x)
{
    int64_t lo = 0, hi = R->n_km;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (R->km_key[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

/* ---- seeding: all MEMs >= k on both strands (first bullet) ---- */
int qmo_collect_seeds(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *read, int len, qmo_seed_t *S)
{
    const int k = R->k;
    const int occ_cap = o->max_occ < QMO_OCC_CAP ? o->max_occ : QMO_OCC_CAP;
    int n = 0, q, i, pass;
    for (q = 0; q + k <= len; ++q) {
        uint64_t fw = 0, rc = 0;
        int bad = 0;
        for (i = 0; i < k; ++i) {
            int c = read[q + i];
            if (c > 3) { bad = 1; break; }
            fw = (fw << 2) | (uint64_t)c;
            rc |= (uint64_t)(3 - c) << (2 * i);
        }
        if (bad) continue;
        for (pass = 0; pass < 2; ++pass) {
            uint64_t key = pass ? rc : fw;
            int64_t a = km_lower(R, key), b = a, t;
            while (b < R->n_km && R->km_key[b] == key) ++b;
            if (b == a || b - a > occ_cap) continue;
            for (t = a; t < b; ++t) {
                int64_t p = R->km_pos[t];
                int64_t rpos = pass ? 2 * R->l_pac - p - k : p;
                int m, found = 0;
                for (m = 0; m < n; ++m)
                    if (S[m].rbeg - S[m].qbeg == rpos - q && S[m].qbeg + S[m].len - k + 1 == q) { ++S[m].len; found = 1; break; }
                if (!found && n < QMO_MAX_SEEDS) { S[n].rbeg = rpos; S[n].qbeg = q; S[n].len = k; ++n; }
            }
        }
    }
    /* order: (qbeg, rbeg) ascending */
    for (i = 1; i < n; ++i) {
        qmo_seed_t x = S[i];
        int j = i - 1;
        while (j >= 0 && (S[j].qbeg > x.qbeg || (S[j].qbeg == x.qbeg && S[j].rbeg > x.rbeg))) { S[j + 1] = S[j]; --j; }
        S[j + 1] = x;
    }
    return n;
}

/* ---- chaining (bwamem.c mem_chain / test_and_merge / mem_chain_weight / mem_chain_flt) ---- */
typedef struct {
    int64_t pos;
    int rid, n, w, kept, first;
    int sidx[QMO_MAX_SEEDS];
} chain_t;

static inline int max_gap_for(const qmo_opt_t *o, int qlen)
{
    int l_del = (int)((double)(qlen * o->a - o->o_del) / o->e_del + 1.);
    int l_ins = (int)((double)(qlen * o->a - o->o_ins) / o->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    if (l < 1) l = 1;
    return l < o->w << 1 ? l : o->w << 1;
}

static int seed_rid(const qmo_ref_t *R, const qmo_seed_t *s)
{
    int rev;
    int64_t f = depos(R, s->rbeg, &rev);
    return pos2rid(R, f);
}

static int try_merge(const qmo_ref_t *R, const qmo_opt_t *o, chain_t *c, const qmo_seed_t *S, int si, int rid)
{
    const qmo_seed_t *p = &S[si], *first = &S[c->sidx[0]], *last = &S[c->sidx[c->n - 1]];
    int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len, x, y;
    if (rid != c->rid) return 0;
    if (p->qbeg >= first->qbeg && p->qbeg + p->len <= qend && p->rbeg >= first->rbeg && p->rbeg + p->len <= rend)
        return 1;                                  /* contained: swallowed */
    if ((last->rbeg < R->l_pac || first->rbeg < R->l_pac) && p->rbeg >= R->l_pac) return 0;
    x = p->qbeg - last->qbeg;
    y = p->rbeg - last->rbeg;
    if (y >= 0 && x - y <= o->w && y - x <= o->w && x - last->len < o->max_chain_gap && y - last->len < o->max_chain_gap) {
        c->sidx[c->n++] = si;
        return 1;
    }
    return 0;
}

static int chain_weight(const qmo_seed_t *S, const chain_t *c)
{
    int64_t end;
    int j, w = 0, tmp;
    for (j = 0, end = 0; j < c->n; ++j) {
        const qmo_seed_t *s = &S[c->sidx[j]];
        if (s->qbeg >= end) w += s->len;
        else if (s->qbeg + s->len > end) w += (int)(s->qbeg + s->len - end);
        if (s->qbeg + s->len > end) end = s->qbeg + s->len;
    }
    tmp = w; w = 0;
    for (j = 0, end = 0; j < c->n; ++j) {
        const qmo_seed_t *s = &S[c->sidx[j]];
        if (s->rbeg >= end) w += s->len;
        else if (s->rbeg + s->len > end) w += (int)(s->rbeg + s->len - end);
        if (s->rbeg + s->len > end) end = s->rbeg + s->len;
    }
    return w < tmp ? w : tmp;
}

/* builds chains in chn[], returns the number kept after filtering, in processing order */
static int build_chains(const qmo_ref_t *R, const qmo_opt_t *o, const qmo_seed_t *S, int n_seeds, chain_t *chn)
{
    int order[QMO_MAX_SEEDS];        /* chain indices sorted by pos (the B-tree) */
    int n_chn = 0, i, k;
    for (i = 0; i < n_seeds; ++i) {
        int rid = seed_rid(R, &S[i]);
        int lower = -1, at;
        for (k = 0; k < n_chn; ++k) { if (chn[order[k]].pos <= S[i].rbeg) lower = k; else break; }
        if (lower >= 0 && try_merge(R, o, &chn[order[lower]], S, i, rid)) continue;
        chn[n_chn].pos = S[i].rbeg; chn[n_chn].rid = rid; chn[n_chn].n = 1; chn[n_chn].sidx[0] = i;
        chn[n_chn].w = 0; chn[n_chn].kept = 0; chn[n_chn].first = -1;
        at = lower + 1;                                   /* after every chain with pos <= rbeg */
        memmove(order + at + 1, order + at, sizeof(int) * (n_chn - at));
        order[at] = n_chn++;
    }
    if (n_chn == 0) return 0;
    /* filter (mem_chain_flt): weight, stable sort by weight descending, overlap test */
    chain_t tmp[QMO_MAX_SEEDS];
    int m = 0;
    for (k = 0; k < n_chn; ++k) {
        chain_t *c = &chn[order[k]];
        c->w = chain_weight(S, c);
        if (c->w >= o->min_chain_weight) tmp[m++] = *c;
    }
    n_chn = m;
    for (i = 1; i < n_chn; ++i) {
        chain_t x = tmp[i];
        int j = i - 1;
        while (j >= 0 && tmp[j].w < x.w) { tmp[j + 1] = tmp[j]; --j; }
        tmp[j + 1] = x;
    }
    if (n_chn == 0) return 0;
    int kept_idx[QMO_MAX_SEEDS], n_kept = 0;
#define CBEG(c) (S[(c).sidx[0]].qbeg)
#define CEND(c) (S[(c).sidx[(c).n - 1]].qbeg + S[(c).sidx[(c).n - 1]].len)
    tmp[0].kept = 3; kept_idx[n_kept++] = 0;
    for (i = 1; i < n_chn; ++i) {
        int large_ovlp = 0;
        for (k = 0; k < n_kept; ++k) {
            int j = kept_idx[k];
            int b_max = CBEG(tmp[j]) > CBEG(tmp[i]) ? CBEG(tmp[j]) : CBEG(tmp[i]);
            int e_min = CEND(tmp[j]) < CEND(tmp[i]) ? CEND(tmp[j]) : CEND(tmp[i]);
            if (e_min > b_max) {
                int li = CEND(tmp[i]) - CBEG(tmp[i]), lj = CEND(tmp[j]) - CBEG(tmp[j]);
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * o->mask_level && min_l < o->max_chain_gap) {
                    large_ovlp = 1;
                    if (tmp[j].first < 0) tmp[j].first = i;
                    if (tmp[i].w < tmp[j].w * o->drop_ratio && tmp[j].w - tmp[i].w >= o->min_seed_len << 1) break;
                }
            }
        }
        if (k == n_kept) { kept_idx[n_kept++] = i; tmp[i].kept = large_ovlp ? 2 : 3; }
    }
    for (i = 0; i < n_kept; ++i) { chain_t *c = &tmp[kept_idx[i]]; if (c->first >= 0) tmp[c->first].kept = 1; }
    /* max_chain_extend = 2^30: the truncation loop of mem_chain_flt never fires */
    m = 0;
    for (i = 0; i < n_chn; ++i) if (tmp[i].kept) chn[m++] = tmp[i];
#undef CBEG
#undef CEND
    return m;
}

/* ---- extension log ---- */
static void log_push(qmo_ext_log_t *L, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int h0, int w,
                     int end_bonus, uint32_t flags, const qmo_ext_t *res, int w_used, int64_t cells)
{
    if (!L) return;
    if (L->seq_len + qlen + tlen > L->seq_cap) {
        L->seq_cap = (L->seq_len + qlen + tlen) * 2 + 1024;
        L->seq = (uint8_t *)realloc(L->seq, L->seq_cap);
    }
    if (L->n == L->cap) {
        L->cap = L->cap * 2 + 1024;
        L->tasks = (qmo_ext_task_t *)realloc(L->tasks, sizeof(qmo_ext_task_t) * L->cap);
        L->results = (qmo_ext_t *)realloc(L->results, sizeof(qmo_ext_t) * L->cap);
        L->w_used = (int32_t *)realloc(L->w_used, 4 * L->cap);
        L->cells = (int64_t *)realloc(L->cells, 8 * L->cap);
    }
    qmo_ext_task_t *t_ = &L->tasks[L->n];
    t_->q_off = (uint32_t)L->seq_len; memcpy(L->seq + L->seq_len, q, qlen); L->seq_len += qlen;
    t_->t_off = (uint32_t)L->seq_len; memcpy(L->seq + L->seq_len, t, tlen); L->seq_len += tlen;
    t_->qlen = qlen; t_->tlen = tlen; t_->h0 = h0; t_->w = w; t_->end_bonus = end_bonus; t_->flags = flags;
    L->results[L->n] = *res; L->w_used[L->n] = w_used; L->cells[L->n] = cells;
    ++L->n;
}

/* one extension with bwa's band retry (MAX_BAND_TRY = 2); prev0 = initial "previous score" */
static int extend_retry(const qmo_opt_t *o, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int h0,
                        int end_bonus, int prev0, qmo_ext_t *res, int *w_used, qmo_ext_log_t *L, int64_t *cells_total)
{
    int i, prev = prev0;
    int64_t cells = 0;
    for (i = 0; i < 2; ++i) {
        *w_used = o->w << i;
        cells += qmo_ksw_extend2(qlen, q, tlen, t, o, *w_used, end_bonus, h0, res);
        if (res->score == prev || res->max_off < (*w_used >> 1) + (*w_used >> 2)) break;
        prev = res->score;
    }
    log_push(L, qlen, q, tlen, t, h0, o->w, end_bonus, 1u | (prev0 >= 0 ? 2u : 0u), res, *w_used, cells);
    if (cells_total) *cells_total += cells;
    return res->score;
}

/* ---- mem_chain2aln ---- */
static void chain_to_regs(const qmo_ref_t *R, const qmo_opt_t *o, int l_query, const uint8_t *query,
                          const qmo_seed_t *S, const chain_t *c, qmo_reg_t *av, int *n_av,
                          qmo_ext_log_t *L, int64_t *cells_total)
{
    const int64_t l_pac = R->l_pac;
    int64_t rmax0 = l_pac << 1, rmax1 = 0;
    int i, k;
    if (c->n == 0) return;
    for (i = 0; i < c->n; ++i) {
        const qmo_seed_t *t = &S[c->sidx[i]];
        int64_t b = t->rbeg - (t->qbeg + max_gap_for(o, t->qbeg));
        int tail = l_query - t->qbeg - t->len;
        int64_t e = t->rbeg + t->len + (tail + max_gap_for(o, tail));
        if (b < rmax0) rmax0 = b;
        if (e > rmax1) rmax1 = e;
    }
    if (rmax0 < 0) rmax0 = 0;
    if (rmax1 > l_pac << 1) rmax1 = l_pac << 1;
    if (rmax0 < l_pac && l_pac < rmax1) { if (S[c->sidx[0]].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
    {   /* bns_fetch_seq: clamp to the contig holding the first seed */
        int rev;
        int64_t f = depos(R, S[c->sidx[0]].rbeg, &rev);
        int rid = pos2rid(R, f);
        int64_t far_beg = R->off[rid], far_end = R->off[rid] + R->len[rid];
        if (rev) { int64_t t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
        if (rmax0 < far_beg) rmax0 = far_beg;
        if (rmax1 > far_end) rmax1 = far_end;
    }
    /* seeds by descending length, ties by descending index */
    uint64_t srt[QMO_MAX_SEEDS];
    for (i = 0; i < c->n; ++i) srt[i] = (uint64_t)S[c->sidx[i]].len << 32 | (uint32_t)i;
    for (i = 1; i < c->n; ++i) { uint64_t x = srt[i]; int j = i - 1; while (j >= 0 && srt[j] > x) { srt[j + 1] = srt[j]; --j; } srt[j + 1] = x; }

    uint8_t *qs = (uint8_t *)malloc(l_query + 1);
    uint8_t *rs = (uint8_t *)malloc((size_t)(rmax1 - rmax0) + 1);
    for (k = c->n - 1; k >= 0; --k) {
        const qmo_seed_t *s = &S[c->sidx[(uint32_t)srt[k]]];
        for (i = 0; i < *n_av; ++i) {
            const qmo_reg_t *p = &av[i];
            int64_t rd;
            int qd, w, mg;
            if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
            if (s->len - p->seedlen0 > .1 * l_query) continue;
            qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
            mg = max_gap_for(o, qd < rd ? qd : (int)rd);
            w = mg < p->w ? mg : p->w;
            if (qd - rd < w && rd - qd < w) break;
            qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
            mg = max_gap_for(o, qd < rd ? qd : (int)rd);
            w = mg < p->w ? mg : p->w;
            if (qd - rd < w && rd - qd < w) break;
        }
        if (i < *n_av) {
            for (i = k + 1; i < c->n; ++i) {
                const qmo_seed_t *t;
                if (srt[i] == 0) continue;
                t = &S[c->sidx[(uint32_t)srt[i]]];
                if (t->len < s->len * .95) continue;
                if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
            }
            if (i == c->n) { srt[k] = 0; continue; }
        }
        if (*n_av >= QMO_MAX_REGS) { srt[k] = 0; continue; }     /* documented hard cap */
        qmo_reg_t *a = &av[(*n_av)++];
        int aw0 = o->w, aw1 = o->w;
        memset(a, 0, sizeof(*a));
        a->w = o->w; a->score = a->truesc = -1; a->rid = c->rid; a->secondary = -1;
        if (s->qbeg) {
            qmo_ext_t r;
            int64_t tlen = s->rbeg - rmax0;
            for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
            for (i = 0; i < tlen; ++i) rs[i] = (uint8_t)ref_base(R, s->rbeg - 1 - i);
            a->score = extend_retry(o, s->qbeg, qs, (int)tlen, rs, s->len * o->a, o->pen_clip5, -1, &r, &aw0, L, cells_total);
            if (r.gscore <= 0 || r.gscore <= a->score - o->pen_clip5) { a->qb = s->qbeg - r.qle; a->rb = s->rbeg - r.tle; a->truesc = a->score; }
            else { a->qb = 0; a->rb = s->rbeg - r.gtle; a->truesc = r.gscore; }
        } else { a->score = a->truesc = s->len * o->a; a->qb = 0; a->rb = s->rbeg; }
        if (s->qbeg + s->len != l_query) {
            qmo_ext_t r;
            int sc0 = a->score, qe = s->qbeg + s->len;
            int64_t re = s->rbeg + s->len, tlen = rmax1 - re;
            for (i = 0; i < tlen; ++i) rs[i] = (uint8_t)ref_base(R, re + i);
            a->score = extend_retry(o, l_query - qe, query + qe, (int)tlen, rs, sc0, o->pen_clip3, sc0, &r, &aw1, L, cells_total);
            if (r.gscore <= 0 || r.gscore <= a->score - o->pen_clip3) { a->qe = qe + r.qle; a->re = re + r.tle; a->truesc += a->score - sc0; }
            else { a->qe = l_query; a->re = re + r.gtle; a->truesc += r.gscore - sc0; }
        } else { a->qe = l_query; a->re = s->rbeg + s->len; }
        for (i = 0, a->seedcov = 0; i < c->n; ++i) {
            const qmo_seed_t *t = &S[c->sidx[i]];
            if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) a->seedcov += t->len;
        }
        a->w = aw0 > aw1 ? aw0 : aw1;
        a->seedlen0 = s->len;
    }
    free(qs); free(rs);
}

/* ---- mem_sort_dedup_patch; stable sorts.  Patching (bwamem.c mem_patch_reg: two collinear hits of a read that one banded global
 * alignment explains at >= 90 % of the score their lengths predict become one) runs when the caller hands in the read (R, query):
 * mem_align1_core does, mem_matesw does not.  Off unless o->flags & QMO_F_PATCH (a deviation the product shares by default). ---- */
#define PATCH_MAX_R_BW 0.05f
#define PATCH_MIN_SC_RATIO 0.90f
#define QMO_PATCH_MAX_QUERY 256      /* longer spans are not patched (limit shared with the product's scalar DP) */
static int gen_cigar(const qmo_ref_t *R, const qmo_opt_t *o, int w_, int l_query, const uint8_t *query,
                     int64_t rb, int64_t re, int *n_cigar, uint32_t *cigar, int *nm);
long long g_qmo_patch_tried = 0, g_qmo_patch_done = 0;      /* diagnostics */

static int patch_reg(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *query, const qmo_reg_t *a, const qmo_reg_t *b, int *_w)
{
    int w, score, q_s, r_s, n_cigar = 0, nm;
    uint32_t cig[QMO_MAX_CIGAR];
    double r;
    if (a->rb < R->l_pac && b->rb >= R->l_pac) return 0;                          /* on different strands */
    if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;             /* not collinear */
    w = (int)((a->re - b->rb) - (a->qe - b->qb));                                 /* required bandwidth */
    w = w > 0 ? w : -w;
    r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);   /* relative bandwidth */
    r = r > 0. ? r : -r;
    if (a->re < b->rb || a->qe < b->qb) {                                         /* no overlap on query or on reference */
        if (w > o->w << 1 || r >= PATCH_MAX_R_BW) return 0;
    } else if (w > o->w << 2 || r >= PATCH_MAX_R_BW * 2) return 0;                /* more permissive if overlapping on both */
    if (b->qe - a->qb > QMO_PATCH_MAX_QUERY || b->re - a->rb > 2 * QMO_PATCH_MAX_QUERY) return 0;
    w += a->w + b->w;
    w = w < o->w << 2 ? w : o->w << 2;
#pragma omp atomic
    ++g_qmo_patch_tried;
    score = gen_cigar(R, o, w, b->qe - a->qb, query + a->qb, a->rb, b->re, &n_cigar, cig, &nm);
    q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);
    r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);
    if ((double)score / (q_s > r_s ? q_s : r_s) < PATCH_MIN_SC_RATIO) return 0;
    *_w = w;
#pragma omp atomic
    ++g_qmo_patch_done;
    return score;
}

static int sort_dedup_patch(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *query, int n, qmo_reg_t *a);
static int sort_dedup(const qmo_opt_t *o, int n, qmo_reg_t *a) { return sort_dedup_patch(0, o, 0, n, a); }

static int sort_dedup_patch(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *query, int n, qmo_reg_t *a)
{
    int i, j, m;
    if (n <= 1) return n;
    for (i = 1; i < n; ++i) { qmo_reg_t x = a[i]; j = i - 1; while (j >= 0 && a[j].re > x.re) { a[j + 1] = a[j]; --j; } a[j + 1] = x; }
    for (i = 1; i < n; ++i) {
        qmo_reg_t *p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + o->max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + o->max_chain_gap; --j) {
            qmo_reg_t *q = &a[j];
            int64_t orr, oq, mr, mq;
            if (q->qe == q->qb) continue;
            orr = q->re - p->rb;
            oq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
            mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if (orr > o->mask_level_redun * mr && oq > o->mask_level_redun * mq) {
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (query && (o->flags & QMO_F_PATCH) && q->rb < p->rb) {
                int w, score = patch_reg(R, o, query, q, p, &w);
                if (score > 0) {                                  /* merge q into p (n_comp is not kept here) */
                    p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
                    p->sub = p->sub > q->sub ? p->sub : q->sub;
                    p->csub = p->csub > q->csub ? p->csub : q->csub;
                    p->qb = q->qb; p->rb = q->rb;
                    p->truesc = p->score = score;
                    p->w = w;
                    q->qe = q->qb;
                }
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i) if (a[i].qe > a[i].qb) a[m++] = a[i];
    n = m;
    for (i = 1; i < n; ++i) {
        qmo_reg_t x = a[i]; j = i - 1;
        while (j >= 0 && !(a[j].score > x.score || (a[j].score == x.score && (a[j].rb < x.rb || (a[j].rb == x.rb && a[j].qb <= x.qb))))) { a[j + 1] = a[j]; --j; }
        a[j + 1] = x;
    }
    for (i = 1; i < n; ++i)
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i) if (a[i].qe > a[i].qb) a[m++] = a[i];
    return n ? m : 0;
}

/* bwa's own seeds (opt.flags & QMO_F_FM_SEEDS): SMEMs + re-seeding + third round through the FM-index attached with
 * qmo_ref_set_fm, in the order mem_chain visits them, the first QMO_MAX_SEEDS of them */
void qmo_ref_set_fm(qmo_ref_t *R, const void *fm, int max_mem_intv) { R->fm = fm; R->fm_max_mem_intv = max_mem_intv; }
static int fm_collect(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *q, int len, qmo_seed_t *S)
{
    int64_t tmp[3 * QMO_MAX_SEEDS];
    int i, n;
    if (!R->fm) return 0;
    n = qmo_fm_seeds(R->fm, R, o, len, q, R->fm_max_mem_intv, tmp, QMO_MAX_SEEDS);
    if (n > QMO_MAX_SEEDS) n = QMO_MAX_SEEDS;
    for (i = 0; i < n; ++i) { S[i].rbeg = tmp[3 * i]; S[i].qbeg = (int32_t)tmp[3 * i + 1]; S[i].len = (int32_t)tmp[3 * i + 2]; }
    return n;
}

void qmo_align_se(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n, const uint8_t *reads, int stride,
                  const int32_t *lens, qmo_seed_t *seeds, int32_t *n_seeds, qmo_reg_t *regs, int32_t *n_regs,
                  qmo_ext_log_t *log, int64_t *cells_total)
{
    int64_t r, cells = 0;
    /* serial when a log is requested (order matters); otherwise parallel over reads */
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cells) if (log == 0)
    for (r = 0; r < n; ++r) {
        qmo_seed_t S[QMO_MAX_SEEDS];
        chain_t *chn = (chain_t *)malloc(sizeof(chain_t) * QMO_MAX_SEEDS);
        qmo_reg_t av[QMO_MAX_REGS];
        const uint8_t *q = reads + r * stride;
        int ns = (o->flags & QMO_F_FM_SEEDS) ? fm_collect(R, o, q, lens[r], S) : qmo_collect_seeds(R, o, q, lens[r], S);
        int nc = build_chains(R, o, S, ns, chn), c, nav = 0;
        int64_t mycells = 0;
        for (c = 0; c < nc; ++c) chain_to_regs(R, o, lens[r], q, S, &chn[c], av, &nav, log, &mycells);
        nav = sort_dedup_patch(R, o, q, nav, av);
        if (seeds) { memcpy(seeds + r * QMO_MAX_SEEDS, S, sizeof(qmo_seed_t) * ns); n_seeds[r] = ns; }
        if (regs) { memcpy(regs + r * QMO_MAX_REGS, av, sizeof(qmo_reg_t) * nav); n_regs[r] = nav; }
        cells += mycells;
        free(chn);
    }
    if (cells_total) *cells_total = cells;
}

/* ---- pairing (bwamem_pair.c) ---- */
static inline int infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
    int r1 = (b1 >= l_pac), r2 = (b2 >= l_pac);
    int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;
    *dist = p2 > b1 ? p2 - b1 : b1 - p2;
    return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

static int cal_sub(const qmo_opt_t *o, const qmo_reg_t *a, int n)
{
    int j;
    for (j = 1; j < n; ++j) {
        int b_max = a[j].qb > a[0].qb ? a[j].qb : a[0].qb;
        int e_min = a[j].qe < a[0].qe ? a[j].qe : a[0].qe;
        if (e_min > b_max) {
            int l0 = a[0].qe - a[0].qb, lj = a[j].qe - a[j].qb;
            int min_l = lj < l0 ? lj : l0;
            if (e_min - b_max >= min_l * o->mask_level) break;
        }
    }
    return j < n ? a[j].score : o->min_seed_len * o->a;
}

static int u64_cmp(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

void qmo_pestat(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, const qmo_reg_t *regs,
                const int32_t *n_regs, qmo_pestat_t pes[4])
{
    uint64_t *isz[4];
    int64_t cnt[4] = {0, 0, 0, 0}, i, max = 0;
    int d;
    memset(pes, 0, 4 * sizeof(qmo_pestat_t));
    for (d = 0; d < 4; ++d) isz[d] = (uint64_t *)malloc(8 * (size_t)(n_pairs + 1));
    for (i = 0; i < n_pairs; ++i) {
        const qmo_reg_t *r0 = regs + (2 * i) * QMO_MAX_REGS, *r1 = regs + (2 * i + 1) * QMO_MAX_REGS;
        int n0 = n_regs[2 * i], n1 = n_regs[2 * i + 1], dir;
        int64_t is;
        if (n0 == 0 || n1 == 0) continue;
        if (cal_sub(o, r0, n0) > 0.8 * r0[0].score) continue;
        if (cal_sub(o, r1, n1) > 0.8 * r1[0].score) continue;
        if (r0[0].rid != r1[0].rid) continue;
        dir = infer_dir(R->l_pac, r0[0].rb, r1[0].rb, &is);
        if (is && is <= o->max_ins) isz[dir][cnt[dir]++] = (uint64_t)is;
    }
    for (d = 0; d < 4; ++d) {
        qmo_pestat_t *r = &pes[d];
        uint64_t *q = isz[d];
        int64_t n = cnt[d], x;
        int p25, p50, p75;
        if (n < 10) { r->failed = 1; continue; }
        qsort(q, n, 8, u64_cmp);
        p25 = (int)q[(int)(.25 * n + .499)];
        p50 = (int)q[(int)(.50 * n + .499)];
        p75 = (int)q[(int)(.75 * n + .499)];
        (void)p50;
        r->low = (int)(p25 - 2.0 * (p75 - p25) + .499);
        if (r->low < 1) r->low = 1;
        r->high = (int)(p75 + 2.0 * (p75 - p25) + .499);
        for (i = x = 0, r->avg = 0; i < n; ++i) if ((int64_t)q[i] >= r->low && (int64_t)q[i] <= r->high) { r->avg += q[i]; ++x; }
        r->avg /= x;
        for (i = 0, r->std = 0; i < n; ++i) if ((int64_t)q[i] >= r->low && (int64_t)q[i] <= r->high) r->std += (q[i] - r->avg) * (q[i] - r->avg);
        r->std = sqrt(r->std / x);
        r->low = (int)(p25 - 3.0 * (p75 - p25) + .499);
        r->high = (int)(p75 + 3.0 * (p75 - p25) + .499);
        if (r->low > r->avg - 4.0 * r->std) r->low = (int)(r->avg - 4.0 * r->std + .499);
        if (r->high < r->avg + 4.0 * r->std) r->high = (int)(r->avg + 4.0 * r->std + .499);
        if (r->low < 1) r->low = 1;
    }
    for (d = 0; d < 4; ++d) if (cnt[d] > max) max = cnt[d];
    for (d = 0; d < 4; ++d) if (!pes[d].failed && cnt[d] < max * 0.05) pes[d].failed = 1;
    for (d = 0; d < 4; ++d) free(isz[d]);
}

static inline uint64_t hash64(uint64_t key)
{
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3);   key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

/* mem_mark_primary_se: sort by (score desc, hash asc), mark secondaries, fill sub / sub_n */
static void mark_primary(const qmo_opt_t *o, int n, qmo_reg_t *a, uint64_t id)
{
    uint64_t hsh[QMO_MAX_REGS];
    int z[QMO_MAX_REGS], nz = 0, i, k, tmp;
    if (n == 0) return;
    for (i = 0; i < n; ++i) { a[i].sub = 0; a[i].sub_n = 0; a[i].secondary = -1; hsh[i] = hash64(id + i); }
    for (i = 1; i < n; ++i) {
        qmo_reg_t x = a[i]; uint64_t hx = hsh[i]; int j = i - 1;
        while (j >= 0 && !(a[j].score > x.score || (a[j].score == x.score && hsh[j] <= hx))) { a[j + 1] = a[j]; hsh[j + 1] = hsh[j]; --j; }
        a[j + 1] = x; hsh[j + 1] = hx;
    }
    tmp = o->a + o->b;
    if (o->o_del + o->e_del > tmp) tmp = o->o_del + o->e_del;
    if (o->o_ins + o->e_ins > tmp) tmp = o->o_ins + o->e_ins;
    z[nz++] = 0;
    for (i = 1; i < n; ++i) {
        for (k = 0; k < nz; ++k) {
            int j = z[k];
            int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {
                int li = a[i].qe - a[i].qb, lj = a[j].qe - a[j].qb;
                int min_l = li < lj ? li : lj;
                if (e_min - b_max >= min_l * o->mask_level) {
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i; else a[i].secondary = z[k];
    }
}

static int approx_mapq(const qmo_opt_t *o, const qmo_reg_t *a)
{
    int mapq, l, sub = a->sub ? a->sub : o->min_seed_len * o->a;
    double identity, tmp;
    if (a->csub > sub) sub = a->csub;
    if (sub >= a->score) return 0;
    l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
    identity = 1. - (double)(l * o->a - a->score) / (o->a + o->b) / l;
    if (a->score == 0) mapq = 0;
    else {
        tmp = l < o->mapq_coef_len ? 1. : log((double)o->mapq_coef_len) / log((double)l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a->score - sub) / o->a * tmp * tmp + .499);
    }
    if (a->sub_n > 0) mapq -= (int)(4.343 * log((double)(a->sub_n + 1)) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    return mapq;                                   /* frac_rep = 0 */
}

static inline int raw_mapq(int diff, int a) { return (int)(6.02 * diff / a + .499); }

typedef struct { uint64_t x, y; } p128_t;
static void sort128(p128_t *v, int n)
{
    int i;
    for (i = 1; i < n; ++i) { p128_t t = v[i]; int j = i - 1; while (j >= 0 && (v[j].x > t.x || (v[j].x == t.x && v[j].y > t.y))) { v[j + 1] = v[j]; --j; } v[j + 1] = t; }
}

/* mem_pair: returns best pair score o (0 = none); z[] = chosen region per end */
static int pair_up(const qmo_ref_t *R, const qmo_opt_t *o, const qmo_pestat_t pes[4], qmo_reg_t *a[2], const int n_pri[2],
                   uint64_t id, int *sub, int *n_sub, int z[2])
{
    p128_t v[2 * QMO_MAX_REGS], u[8 * QMO_MAX_REGS * QMO_MAX_REGS];
    int nv = 0, nu = 0, r, i, k, y[4], ret;
    const int64_t l_pac = R->l_pac;
    for (r = 0; r < 2; ++r)
        for (i = 0; i < n_pri[r]; ++i) {
            const qmo_reg_t *e = &a[r][i];
            uint64_t fx = e->rb < l_pac ? (uint64_t)e->rb : (uint64_t)((l_pac << 1) - 1 - e->rb);
            v[nv].x = (uint64_t)e->rid << 32 | (fx - (uint64_t)R->off[e->rid]);
            v[nv].y = (uint64_t)e->score << 32 | (uint64_t)(i << 2) | (uint64_t)((e->rb >= l_pac) << 1) | (uint64_t)r;
            ++nv;
        }
    sort128(v, nv);
    y[0] = y[1] = y[2] = y[3] = -1;
    for (i = 0; i < nv; ++i) {
        for (r = 0; r < 2; ++r) {
            int dir = r << 1 | (int)(v[i].y >> 1 & 1), which;
            if (pes[dir].failed) continue;
            which = r << 1 | ((int)(v[i].y & 1) ^ 1);
            if (y[which] < 0) continue;
            for (k = y[which]; k >= 0; --k) {
                int64_t dist;
                int q;
                double ns;
                if ((int)(v[k].y & 3) != which) continue;
                dist = (int64_t)v[i].x - (int64_t)v[k].x;
                if (dist > pes[dir].high) break;
                if (dist < pes[dir].low) continue;
                ns = (dist - pes[dir].avg) / pes[dir].std;
                q = (int)((v[i].y >> 32) + (v[k].y >> 32) + .721 * log(2. * erfc(fabs(ns) * M_SQRT1_2)) * o->a + .499);
                if (q < 0) q = 0;
                u[nu].y = (uint64_t)k << 32 | (uint64_t)i;
                u[nu].x = (uint64_t)q << 32 | (hash64(u[nu].y ^ id << 8) & 0xffffffffU);
                ++nu;
            }
        }
        y[v[i].y & 3] = i;
    }
    if (nu) {
        int tmp = o->a + o->b;
        if (o->o_del + o->e_del > tmp) tmp = o->o_del + o->e_del;
        if (o->o_ins + o->e_ins > tmp) tmp = o->o_ins + o->e_ins;
        sort128(u, nu);
        i = (int)(u[nu - 1].y >> 32); k = (int)(u[nu - 1].y << 32 >> 32);
        z[v[i].y & 1] = (int)(v[i].y << 32 >> 34);
        z[v[k].y & 1] = (int)(v[k].y << 32 >> 34);
        ret = (int)(u[nu - 1].x >> 32);
        *sub = nu > 1 ? (int)(u[nu - 2].x >> 32) : 0;
        for (i = nu - 2, *n_sub = 0; i >= 0; --i) if (*sub - (int)(u[i].x >> 32) <= tmp) ++*n_sub;
    } else { ret = 0; *sub = 0; *n_sub = 0; }
    return ret;
}

/* ---- CIGAR (bwamem.c mem_reg2aln, bwa.c bwa_gen_cigar2) ---- */
static inline int infer_bw(int l1, int l2, int score, int a, int q, int r)
{
    int w;
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    w = (int)(((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.));
    if (w < abs(l1 - l2)) w = abs(l1 - l2);
    return w;
}

/* returns score; cigar in forward-strand order with ops M/I/D; NM via *nm */
static int gen_cigar(const qmo_ref_t *R, const qmo_opt_t *o, int w_, int l_query, const uint8_t *query,
                     int64_t rb, int64_t re, int *n_cigar, uint32_t *cigar, int *nm)
{
    const int64_t l_pac = R->l_pac;
    int rlen = (int)(re - rb), i, score = 0;
    *n_cigar = 0; *nm = -1;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return 0;
    uint8_t *rs = (uint8_t *)malloc(rlen + 1), *qs = (uint8_t *)malloc(l_query + 1);
    for (i = 0; i < rlen; ++i) rs[i] = (uint8_t)ref_base(R, rb + i);
    memcpy(qs, query, l_query);
    if (rb >= l_pac) {            /* reverse both so that gaps are left-aligned on the forward strand */
        for (i = 0; i < l_query >> 1; ++i) { uint8_t t = qs[i]; qs[i] = qs[l_query - 1 - i]; qs[l_query - 1 - i] = t; }
        for (i = 0; i < rlen >> 1; ++i) { uint8_t t = rs[i]; rs[i] = rs[rlen - 1 - i]; rs[rlen - 1 - i] = t; }
    }
    if (l_query == rlen && w_ == 0) {
        cigar[0] = (uint32_t)l_query << 4; *n_cigar = 1;
        for (i = 0; i < l_query; ++i) score += (rs[i] > 3 || qs[i] > 3) ? -1 : (rs[i] == qs[i] ? o->a : -o->b);
    } else {
        int w, max_gap, max_ins, max_del, min_w;
        max_ins = (int)((double)(((l_query + 1) >> 1) * o->a - o->o_ins) / o->e_ins + 1.);
        max_del = (int)((double)(((l_query + 1) >> 1) * o->a - o->o_del) / o->e_del + 1.);
        max_gap = max_ins > max_del ? max_ins : max_del;
        if (max_gap < 1) max_gap = 1;
        w = (max_gap + abs(rlen - l_query) + 1) >> 1;
        if (w > w_) w = w_;
        min_w = abs(rlen - l_query) + 3;
        if (w < min_w) w = min_w;
        score = qmo_ksw_global2(l_query, qs, rlen, rs, o, w, n_cigar, cigar, QMO_MAX_CIGAR - 2);
    }
    if (*n_cigar > 0) {
        int k, x = 0, y = 0, n_mm = 0, n_gap = 0;
        for (k = 0; k < *n_cigar; ++k) {
            int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
            if (op == 0) { for (i = 0; i < len; ++i) if (qs[x + i] != rs[y + i]) ++n_mm; x += len; y += len; }
            else if (op == 2) { if (k > 0 && k < *n_cigar - 1) n_gap += len; y += len; }
            else if (op == 1) { x += len; n_gap += len; }
        }
        *nm = n_mm + n_gap;
    }
    free(rs); free(qs);
    return score;
}

static void reg_to_aln(const qmo_ref_t *R, const qmo_opt_t *o, int l_query, const uint8_t *query,
                       const qmo_reg_t *ar, qmo_aln_t *a)
{
    memset(a, 0, sizeof(*a));
    if (ar == 0 || ar->rb < 0 || ar->re < 0) { a->rid = -1; a->pos = -1; a->flag |= 0x4; return; }
    int qb = ar->qb, qe = ar->qe, i, w2, tmp, nm = -1, score = 0, last_sc = -(1 << 30), is_rev, n_cigar = 0;
    int64_t rb = ar->rb, re = ar->re, pos;
    uint32_t cig[QMO_MAX_CIGAR];
    a->mapq = ar->secondary < 0 ? (uint8_t)approx_mapq(o, ar) : 0;
    if (ar->secondary >= 0) a->flag |= 0x100;
    tmp = infer_bw(qe - qb, (int)(re - rb), ar->truesc, o->a, o->o_del, o->e_del);
    w2 = infer_bw(qe - qb, (int)(re - rb), ar->truesc, o->a, o->o_ins, o->e_ins);
    if (tmp > w2) w2 = tmp;
    if (w2 > o->w) w2 = w2 < ar->w ? w2 : ar->w;
    i = 0;
    do {
        if (w2 > o->w << 2) w2 = o->w << 2;
        score = gen_cigar(R, o, w2, qe - qb, query + qb, rb, re, &n_cigar, cig, &nm);
        if (score == last_sc || w2 == o->w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < ar->truesc - o->a);
    a->nm = nm;
    pos = depos(R, rb < R->l_pac ? rb : re - 1, &is_rev);
    if (n_cigar < 0) { a->rid = -1; a->pos = -1; a->flag |= 0x4; a->n_cigar = 255; return; }   /* cigar overflow */
    if (n_cigar > 0) {          /* squeeze out a leading or trailing deletion */
        if ((cig[0] & 0xf) == 2) { pos += cig[0] >> 4; --n_cigar; memmove(cig, cig + 1, 4 * n_cigar); }
        else if ((cig[n_cigar - 1] & 0xf) == 2) --n_cigar;
    }
    int m = 0;
    if (qb != 0 || qe != l_query) {
        int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
        if (clip5) a->cigar[m++] = (uint32_t)clip5 << 4 | 4;
        for (i = 0; i < n_cigar; ++i) a->cigar[m++] = cig[i];
        if (clip3) a->cigar[m++] = (uint32_t)clip3 << 4 | 4;
    } else for (i = 0; i < n_cigar; ++i) a->cigar[m++] = cig[i];
    a->n_cigar = (uint8_t)m;
    a->rid = pos2rid(R, pos);
    a->pos = (int32_t)(pos - R->off[a->rid]);
    if (is_rev) a->flag |= 0x10;
    a->score = ar->score; a->sub = ar->sub > ar->csub ? ar->sub : ar->csub;
    a->qb = qb; a->qe = qe;
}

static int cigar_rlen(const qmo_aln_t *a)
{
    int k, l = 0;
    for (k = 0; k < a->n_cigar; ++k) { int op = a->cigar[k] & 0xf; if (op == 0 || op == 2) l += a->cigar[k] >> 4; }
    return l;
}

/* flag / mate fields as bwamem.c mem_aln2sam writes them */
static void finish_pair(qmo_aln_t h[2], int extra_flag)
{
    int i;
    int mapped[2] = { h[0].rid >= 0, h[1].rid >= 0 };
    int rev[2] = { (h[0].flag & 0x10) != 0, (h[1].flag & 0x10) != 0 };
    int rlen[2] = { cigar_rlen(&h[0]), cigar_rlen(&h[1]) };
    for (i = 0; i < 2; ++i) {
        qmo_aln_t *p = &h[i], *m = &h[!i];
        p->flag |= 0x1 | (i == 0 ? 0x40 : 0x80) | extra_flag;
        if (!mapped[!i]) p->flag |= 0x8;
        if (mapped[!i] && rev[!i]) p->flag |= 0x20;
        if (!mapped[i] && mapped[!i]) {      /* unmapped read is placed at its mate */
            p->rid = m->rid; p->pos = m->pos;
            if (rev[!i]) p->flag |= 0x10;
        }
        if (!mapped[i] && !mapped[!i]) { p->mate_rid = -1; p->mate_pos = -1; p->tlen = 0; continue; }
        if (mapped[!i]) { p->mate_rid = m->rid; p->mate_pos = m->pos; }
        else { p->mate_rid = p->rid; p->mate_pos = p->pos; if (rev[i]) p->flag |= 0x20; }
        p->tlen = 0;
        if (mapped[0] && mapped[1] && h[0].rid == h[1].rid) {
            int64_t p0 = (i == 0 ? h[0].pos : h[1].pos) + (rev[i] ? rlen[i] - 1 : 0);
            int64_t p1 = (i == 0 ? h[1].pos : h[0].pos) + (rev[!i] ? rlen[!i] - 1 : 0);
            p->tlen = (int32_t)(-(p0 - p1 + (p0 > p1 ? 1 : p0 < p1 ? -1 : 0)));
        }
    }
}

/* ---- mate rescue (bwamem_pair.c mem_matesw, and the loop around it in mem_sam_pe) ----
 * `a` = one region of the anchoring end, ms = the mate as sequenced, ma / *n_ma = the mate's regions (sorted by score).
 * For every orientation with a usable insert-size model in which the mate has no region at a proper distance from
 * `a`, the mate (reverse-complemented when the orientation asks for the opposite strand) is aligned locally inside the
 * window the model implies; a hit of at least min_seed_len joins the mate's list, which is then de-duplicated again.
 * Returns the number of local alignments run.  Limits shared with the product: a window longer than
 * QMO_RESCUE_MAX_WINDOW is not searched; a full list (QMO_MAX_REGS) drops its lowest-scoring entry. */
static int matesw(const qmo_ref_t *R, const qmo_opt_t *o, const qmo_pestat_t pes[4], const qmo_reg_t *a, int l_ms,
                  const uint8_t *ms, qmo_reg_t *ma, int *n_ma, int64_t *cells)
{
    const int64_t l_pac = R->l_pac;
    int skip[4], r, i, n = 0;
    for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
    for (i = 0; i < *n_ma; ++i) {
        int64_t dist;
        r = infer_dir(l_pac, a->rb, ma[i].rb, &dist);
        if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
    }
    if (skip[0] + skip[1] + skip[2] + skip[3] == 4) return 0;
    for (r = 0; r < 4; ++r) {
        int is_rev, is_larger, rid = -1;
        int64_t rb, re;
        if (skip[r]) continue;
        is_rev = (r >> 1) != (r & 1);          /* the mate is searched as its reverse complement */
        is_larger = !(r >> 1);                 /* the mate lies at the larger coordinate          */
        if (!is_rev) {
            rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
            re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
        } else {
            rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
            re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
        }
        if (rb < 0) rb = 0;
        if (re > l_pac << 1) re = l_pac << 1;
        if (rb < re) {                          /* bns_fetch_seq: clip to the contig and strand of the midpoint */
            int mrev;
            int64_t mid = (rb + re) >> 1, far_beg, far_end;
            rid = pos2rid(R, depos(R, mid, &mrev));
            far_beg = R->off[rid]; far_end = far_beg + R->len[rid];
            if (mrev) { int64_t t = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - t; }
            if (rb < far_beg) rb = far_beg;
            if (re > far_end) re = far_end;
        }
        if (a->rid == rid && re - rb >= o->min_seed_len && re - rb <= QMO_RESCUE_MAX_WINDOW) {
            const int tlen = (int)(re - rb);
            uint8_t *ref = (uint8_t *)malloc((size_t)tlen), *seq = (uint8_t *)malloc((size_t)l_ms);
            qmo_sw_t aln;
            for (i = 0; i < tlen; ++i) ref[i] = (uint8_t)ref_base(R, rb + i);
            if (is_rev) for (i = 0; i < l_ms; ++i) seq[l_ms - 1 - i] = ms[i] < 4 ? 3 - ms[i] : 4;
            else memcpy(seq, ms, (size_t)l_ms);
            *cells += qmo_ksw_align2(l_ms, seq, tlen, ref, o, o->min_seed_len * o->a, &aln);
            if (aln.score >= o->min_seed_len && aln.qb >= 0) {
                qmo_reg_t b;
                int at;
                memset(&b, 0, sizeof(b));
                b.rid = a->rid;
                b.qb = is_rev ? l_ms - (aln.qe + 1) : aln.qb;
                b.qe = is_rev ? l_ms - aln.qb : aln.qe + 1;
                b.rb = is_rev ? (l_pac << 1) - (rb + aln.te + 1) : rb + aln.tb;
                b.re = is_rev ? (l_pac << 1) - (rb + aln.tb) : rb + aln.te + 1;
                b.score = aln.score;
                b.csub = aln.score2;
                b.secondary = -1;
                b.seedcov = (int)((b.re - b.rb < b.qe - b.qb ? b.re - b.rb : b.qe - b.qb) >> 1);
                /* insert behind the entries that score at least as much */
                for (at = 0; at < *n_ma; ++at) if (ma[at].score < b.score) break;
                if (*n_ma < QMO_MAX_REGS) ++*n_ma;
                if (at < *n_ma) {
                    for (i = *n_ma - 1; i > at; --i) ma[i] = ma[i - 1];
                    ma[at] = b;
                }
            }
            ++n;
            free(ref); free(seq);
        }
        if (n) *n_ma = sort_dedup(o, *n_ma, ma);
    }
    return n;
}

/* mem_sam_pe's rescue loop: the anchors are the regions of each end within pen_unpaired of its best, taken BEFORE any
 * rescue; end 0's anchors go first.  regs / n_regs are updated in place. */
void qmo_mate_rescue(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, const uint8_t *reads, int stride,
                     const int32_t *lens, qmo_reg_t *regs, int32_t *n_regs, const qmo_pestat_t pes[4],
                     int64_t *n_sw_total, int64_t *cells_total)
{
    int64_t pi, n_sw = 0, cells = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : n_sw, cells)
    for (pi = 0; pi < n_pairs; ++pi) {
        qmo_reg_t *a[2] = { regs + (2 * pi) * QMO_MAX_REGS, regs + (2 * pi + 1) * QMO_MAX_REGS };
        int n[2] = { n_regs[2 * pi], n_regs[2 * pi + 1] }, nb[2] = {0, 0}, i, j;
        qmo_reg_t b[2][QMO_MAX_REGS];
        int64_t mycells = 0;
        for (i = 0; i < 2; ++i)
            for (j = 0; j < n[i]; ++j)
                if (a[i][j].score >= a[i][0].score - o->pen_unpaired) b[i][nb[i]++] = a[i][j];
        for (i = 0; i < 2; ++i)
            for (j = 0; j < nb[i] && j < QMO_MAX_MATESW; ++j)
                n_sw += matesw(R, o, pes, &b[i][j], lens[2 * pi + !i], reads + (2 * pi + !i) * (int64_t)stride, a[!i], &n[!i], &mycells);
        n_regs[2 * pi] = n[0]; n_regs[2 * pi + 1] = n[1];
        cells += mycells;
    }
    if (n_sw_total) *n_sw_total = n_sw;
    if (cells_total) *cells_total = cells;
}

void qmo_pair_and_finish(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, int64_t pair_id0,
                         const uint8_t *reads, int stride, const int32_t *lens,
                         qmo_reg_t *regs, int32_t *n_regs, const qmo_pestat_t pes[4], qmo_aln_t *alns)
{
    int64_t pi;
    if (!(o->flags & QMO_F_NO_RESCUE)) qmo_mate_rescue(R, o, n_pairs, reads, stride, lens, regs, n_regs, pes, 0, 0);
#pragma omp parallel for schedule(dynamic, 256)
    for (pi = 0; pi < n_pairs; ++pi) {
        qmo_reg_t *a[2] = { regs + (2 * pi) * QMO_MAX_REGS, regs + (2 * pi + 1) * QMO_MAX_REGS };
        int n[2] = { n_regs[2 * pi], n_regs[2 * pi + 1] }, n_pri[2], i, z[2] = {0, 0};
        const uint8_t *seq[2] = { reads + (2 * pi) * stride, reads + (2 * pi + 1) * stride };
        int l_seq[2] = { lens[2 * pi], lens[2 * pi + 1] };
        uint64_t id = (uint64_t)(pair_id0 + pi);
        qmo_aln_t h[2];
        int extra_flag = 0, o_sc = 0, subo = 0, n_sub = 0, paired_done = 0;
        mark_primary(o, n[0], a[0], id << 1 | 0);
        mark_primary(o, n[1], a[1], id << 1 | 1);
        n_pri[0] = n[0]; n_pri[1] = n[1];                  /* no ALT contigs */
        if (n_pri[0] && n_pri[1] && (o_sc = pair_up(R, o, pes, a, n_pri, id, &subo, &n_sub, z)) > 0) {
            int is_multi[2], q_pe, score_un, q_se[2], j;
            for (i = 0; i < 2; ++i) {
                for (j = 1; j < n_pri[i]; ++j) if (a[i][j].secondary < 0 && a[i][j].score >= o->T) break;
                is_multi[i] = j < n_pri[i];
            }
            if (!(is_multi[0] || is_multi[1])) {
                score_un = a[0][0].score + a[1][0].score - o->pen_unpaired;
                if (score_un > subo) subo = score_un;
                q_pe = raw_mapq(o_sc - subo, o->a);
                if (n_sub > 0) q_pe -= (int)(4.343 * log((double)(n_sub + 1)) + .499);
                if (q_pe < 0) q_pe = 0;
                if (q_pe > 60) q_pe = 60;
                if (o_sc > score_un) {
                    qmo_reg_t *c[2] = { &a[0][z[0]], &a[1][z[1]] };
                    for (i = 0; i < 2; ++i) {
                        if (c[i]->secondary >= 0) { c[i]->sub = a[i][c[i]->secondary].score; c[i]->secondary = -2; }
                        q_se[i] = approx_mapq(o, c[i]);
                    }
                    for (i = 0; i < 2; ++i) {
                        int cap;
                        q_se[i] = q_se[i] > q_pe ? q_se[i] : q_pe < q_se[i] + 40 ? q_pe : q_se[i] + 40;
                        cap = raw_mapq(c[i]->score - c[i]->csub, o->a);
                        if (q_se[i] > cap) q_se[i] = cap;
                    }
                    extra_flag |= 2;
                } else {
                    z[0] = z[1] = 0;
                    q_se[0] = approx_mapq(o, &a[0][0]);
                    q_se[1] = approx_mapq(o, &a[1][0]);
                }
                for (i = 0; i < 2; ++i) {
                    reg_to_aln(R, o, l_seq[i], seq[i], &a[i][z[i]], &h[i]);
                    h[i].mapq = (uint8_t)q_se[i];
                    h[i].flag &= ~0x100;                 /* the chosen hit is reported as primary */
                }
                paired_done = 1;
            }
        }
        if (!paired_done) {
            for (i = 0; i < 2; ++i) {
                if (n[i] && a[i][0].score >= o->T) reg_to_aln(R, o, l_seq[i], seq[i], &a[i][0], &h[i]);
                else reg_to_aln(R, o, l_seq[i], seq[i], 0, &h[i]);
            }
            if (h[0].rid == h[1].rid && h[0].rid >= 0) {
                int64_t dist;
                int d = infer_dir(R->l_pac, a[0][0].rb, a[1][0].rb, &dist);
                if (!pes[d].failed && dist >= pes[d].low && dist <= pes[d].high) extra_flag |= 2;
            }
        }
        finish_pair(h, extra_flag);
        alns[2 * pi] = h[0]; alns[2 * pi + 1] = h[1];
    }
}
