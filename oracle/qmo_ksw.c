/*
 * qmo_ksw.c -- ORACLE (test infrastructure): restatement of bwa 0.7.17 ksw.c ksw_extend2 and
 * ksw_global2 as specified in SURVEY.md Appendix A.3 / A.4 (upstream source is not vendored in
 * /root/reference; call site rules/bwa.smk:15).  PARITY UNPINNED -- see qmo.h.
 */
#include <stdlib.h>
#include <string.h>
#include "qmo.h"

void qmo_opt_default(qmo_opt_t *o)
{
    memset(o, 0, sizeof(*o));
    o->a = 1; o->b = 4;
    o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1;
    o->w = 100; o->zdrop = 100;
    o->pen_clip5 = o->pen_clip3 = 5;
    o->min_seed_len = 31;           /* rules/bwa.smk:15 passes -k 31 */
    o->max_occ = 500;
    o->T = 30;
    o->pen_unpaired = 17;
    o->max_ins = 10000;
    o->max_chain_gap = 10000;
    o->mapq_coef_len = 50;
    o->mask_level = 0.50f; o->drop_ratio = 0.50f; o->mask_level_redun = 0.95f;
    o->min_chain_weight = 0;
}

static inline int sc(const qmo_opt_t *o, int t, int q)
{   /* 5x5 matrix of bwa_fill_scmat: N against anything = -1 */
    if (t > 3 || q > 3) return -1;
    return t == q ? o->a : -o->b;
}

typedef struct { int32_t h, e; } cell_t;

/* SURVEY.md A.3.  Rows = target, columns = query.  cells[] is reused in place, so columns that a row
 * does not visit keep whatever an earlier row left there. */
int64_t qmo_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                        const qmo_opt_t *o, int w, int end_bonus, int h0, qmo_ext_t *out)
{
    const int gapo_d = o->o_del + o->e_del, gapo_i = o->o_ins + o->e_ins;
    cell_t *cells = (cell_t *)calloc((size_t)qlen + 1, sizeof(cell_t));
    int64_t n_cells = 0;
    int row, col;

    /* row -1: the seed's score decays along a leading insertion */
    cells[0].h = h0;
    if (qlen >= 1) cells[1].h = h0 > gapo_i ? h0 - gapo_i : 0;
    for (col = 2; col <= qlen && cells[col - 1].h > o->e_ins; ++col)
        cells[col].h = cells[col - 1].h - o->e_ins;

    /* shrink the band to what the scores can pay for */
    {
        int best = o->a > -1 ? o->a : -1;          /* max entry of the matrix */
        int lim;
        if (-o->b > best) best = -o->b;
        lim = (int)((double)(qlen * best + end_bonus - o->o_ins) / o->e_ins + 1.);
        if (lim < 1) lim = 1;
        if (w > lim) w = lim;
        lim = (int)((double)(qlen * best + end_bonus - o->o_del) / o->e_del + 1.);
        if (lim < 1) lim = 1;
        if (w > lim) w = lim;
    }

    int best_sc = h0, best_row = -1, best_col = -1, end_row = -1, end_sc = -1, off = 0;
    int lo = 0, hi = qlen;
    for (row = 0; row < tlen; ++row) {
        int f = 0, left, rowmax = 0, rowmax_col = -1;
        const int tb = target[row];
        if (lo < row - w) lo = row - w;
        if (hi > row + w + 1) hi = row + w + 1;
        if (hi > qlen) hi = qlen;
        if (lo == 0) {
            left = h0 - (o->o_del + o->e_del * (row + 1));
            if (left < 0) left = 0;
        } else left = 0;
        for (col = lo; col < hi; ++col) {
            int diag = cells[col].h, e = cells[col].e, h, t;
            cells[col].h = left;
            diag = diag ? diag + sc(o, tb, query[col]) : 0;
            h = diag > e ? diag : e;
            if (f > h) h = f;
            left = h;
            if (!(rowmax > h)) rowmax_col = col;     /* ties: the later column wins */
            if (h > rowmax) rowmax = h;
            t = diag - gapo_d; if (t < 0) t = 0;
            e -= o->e_del; if (t > e) e = t;
            cells[col].e = e;
            t = diag - gapo_i; if (t < 0) t = 0;
            f -= o->e_ins; if (t > f) f = t;
        }
        if (hi > lo) n_cells += hi - lo;
        cells[hi].h = left; cells[hi].e = 0;
        if (col == qlen) {                            /* reached the end of the query */
            if (!(end_sc > left)) end_row = row;     /* ties: the later row wins */
            if (left > end_sc) end_sc = left;
        }
        if (rowmax == 0) break;
        if (rowmax > best_sc) {
            int d = rowmax_col - row;
            best_sc = rowmax; best_row = row; best_col = rowmax_col;
            if (d < 0) d = -d;
            if (d > off) off = d;
        } else if (o->zdrop > 0) {
            int dr = row - best_row, dc = rowmax_col - best_col;
            if (dr > dc) { if (best_sc - rowmax - (dr - dc) * o->e_del > o->zdrop) break; }
            else         { if (best_sc - rowmax - (dc - dr) * o->e_ins > o->zdrop) break; }
        }
        for (col = lo; col < hi && cells[col].h == 0 && cells[col].e == 0; ++col) {}
        lo = col;
        for (col = hi; col >= lo && cells[col].h == 0 && cells[col].e == 0; --col) {}
        hi = col + 2 < qlen ? col + 2 : qlen;
    }
    free(cells);
    out->score = best_sc; out->qle = best_col + 1; out->tle = best_row + 1;
    out->gtle = end_row + 1; out->gscore = end_sc; out->max_off = off;
    return n_cells;
}

#define NEG_INF (-0x40000000)

/* SURVEY.md A.4.  Direction byte per cell: bits 0-1 how H was reached (0 diag, 1 E, 2 F),
 * bit 2 "E extends", bits 4-5 = 2 when "F extends". */
int qmo_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    const qmo_opt_t *o, int w, int *n_cigar, uint32_t *cigar, int max_cigar)
{
    const int gapo_d = o->o_del + o->e_del, gapo_i = o->o_ins + o->e_ins;
    int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    uint8_t *dir = (uint8_t *)malloc((size_t)n_col * (tlen > 0 ? tlen : 1));
    cell_t *cells = (cell_t *)calloc((size_t)qlen + 1, sizeof(cell_t));
    int row, col, score, n = 0, overflow = 0;

    cells[0].h = 0; cells[0].e = NEG_INF;
    for (col = 1; col <= qlen && col <= w; ++col) {
        cells[col].h = -(o->o_ins + o->e_ins * col); cells[col].e = NEG_INF;
    }
    for (; col <= qlen; ++col) cells[col].h = cells[col].e = NEG_INF;

    for (row = 0; row < tlen; ++row) {
        int f = NEG_INF, left, lo, hi;
        uint8_t *drow = dir + (size_t)row * n_col;
        const int tb = target[row];
        lo = row > w ? row - w : 0;
        hi = row + w + 1 < qlen ? row + w + 1 : qlen;
        left = lo == 0 ? -(o->o_del + o->e_del * (row + 1)) : NEG_INF;
        for (col = lo; col < hi; ++col) {
            int m = cells[col].h, e = cells[col].e, h, t;
            uint8_t d;
            cells[col].h = left;
            m += sc(o, tb, query[col]);
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            left = h;
            t = m - gapo_d;
            e -= o->e_del;
            if (e > t) d |= 1 << 2; else e = t;
            cells[col].e = e;
            t = m - gapo_i;
            f -= o->e_ins;
            if (f > t) d |= 2 << 4; else f = t;
            drow[col - lo] = d;
        }
        cells[hi].h = left; cells[hi].e = NEG_INF;
    }
    score = cells[qlen].h;

    if (n_cigar && cigar) {
        /* traceback, emitting run-length ops in reverse, then flip */
        int state = 0, i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
#define PUSH(op_, len_) do { \
            if (n > 0 && (int)(cigar[n - 1] & 0xf) == (op_)) cigar[n - 1] += (uint32_t)(len_) << 4; \
            else if (n < max_cigar) cigar[n++] = (uint32_t)(len_) << 4 | (op_); \
            else overflow = 1; } while (0)
        while (i >= 0 && k >= 0) {
            int lo = i > w ? i - w : 0;
            state = dir[(size_t)i * n_col + (k - lo)] >> (state << 1) & 3;
            if (state == 0)      { PUSH(0, 1); --i; --k; }
            else if (state == 1) { PUSH(2, 1); --i; }
            else                 { PUSH(1, 1); --k; }
        }
        if (i >= 0) PUSH(2, i + 1);
        if (k >= 0) PUSH(1, k + 1);
#undef PUSH
        for (i = 0; i < n >> 1; ++i) { uint32_t t = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = t; }
        *n_cigar = overflow ? -1 : n;
    }
    free(cells); free(dir);
    return score;
}

/* ---- ksw_align2 (ksw.c ksw_u8 / ksw_i16 / ksw_align2): local alignment of the whole query against the target,
 * the kernel of mate rescue (bwamem_pair.c mem_matesw).  Upstream runs Farrar's striped SSE2 loop; its results are
 * those of the plain recurrence below (the striping, the lazy F loop and the zero-score padding columns do not
 * change any value that is read), PROVIDED no 8-bit saturation happens, which the caller guarantees by asking
 * for the byte version only when qlen * a < 250.  What has to be kept exactly are the bookkeeping rules:
 *   - row maximum imax; the best score moves only on a STRICTLY greater row maximum (first row wins), te = that row;
 *   - qe = the smallest query index holding the maximum in row te;
 *   - rows with imax >= minsc are logged in b[]: a new entry unless the last entry's row is i - 1, otherwise the
 *     last entry is overwritten when imax is strictly greater (so a logged row stays where its run peaked first);
 *   - score2 / te2 = best logged entry whose row is further than `score` rows (max = (score + max_sc - 1) / max_sc
 *     with max_sc = a) from te, first such entry on ties;
 *   - the start is found by running the same loop on the reversed prefixes query[0..qe], target[0..te] until the
 *     best score reaches `score` (KSW_XSTOP); tb / qb are only set when that pass ends with exactly `score`.
 * minsc < 0: no sub-optimal score, no start (plain score query).  Returns cells executed. */
static void local_pass(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int q_step, int t_step,
                       const qmo_opt_t *o, int minsc, int endsc, int *score, int *te_out, int *qe_out,
                       int **b_sc, int **b_row, int *n_b_out, int64_t *cells)
{
    const int oe_del = o->o_del + o->e_del, oe_ins = o->o_ins + o->e_ins;
    int *H = (int *)calloc((size_t)qlen + 1, sizeof(int)), *E = (int *)calloc((size_t)qlen + 1, sizeof(int));
    int gmax = 0, te = -1, qe = -1, n_b = 0, m_b = 0, i, j;
    int *bs = 0, *br = 0;
    for (i = 0; i < tlen; ++i) {
        const int tb = target[(int64_t)i * t_step];
        int f = 0, diag = 0, imax = 0, imax_j = -1;      /* H[-1] of every row is 0 */
        for (j = 0; j < qlen; ++j) {
            /* H[j] still holds row i-1; E[j] is the deletion score entering (i, j) */
            int h = diag + sc(o, tb, query[(int64_t)j * q_step]), e = E[j], t;
            diag = H[j];
            if (h < e) h = e;
            if (h < f) h = f;
            if (h < 0) h = 0;
            H[j] = h;
            if (h > imax) { imax = h; imax_j = j; }
            t = h - oe_del; e -= o->e_del; if (e < t) e = t; if (e < 0) e = 0; E[j] = e;
            t = h - oe_ins; f -= o->e_ins; if (f < t) f = t; if (f < 0) f = 0;
        }
        *cells += qlen;
        if (imax >= minsc) {
            if (n_b == 0 || br[n_b - 1] + 1 != i) {
                if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; bs = (int *)realloc(bs, sizeof(int) * m_b); br = (int *)realloc(br, sizeof(int) * m_b); }
                bs[n_b] = imax; br[n_b++] = i;
            } else if (bs[n_b - 1] < imax) { bs[n_b - 1] = imax; br[n_b - 1] = i; }
        }
        if (imax > gmax) {
            gmax = imax; te = i; qe = imax_j;
            if (gmax >= endsc) break;
        }
    }
    free(H); free(E);
    *score = gmax; *te_out = te; *qe_out = qe;
    if (b_sc) { *b_sc = bs; *b_row = br; *n_b_out = n_b; } else { free(bs); free(br); }
}

int64_t qmo_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const qmo_opt_t *o,
                       int minsc, qmo_sw_t *r)
{
    int64_t cells = 0;
    int *bs = 0, *br = 0, n_b = 0, i;
    const int no_limit = 0x10000;
    r->score = 0; r->te = r->qe = -1; r->score2 = -1; r->te2 = -1; r->tb = r->qb = -1;
    local_pass(qlen, query, tlen, target, 1, 1, o, minsc >= 0 ? minsc : no_limit, no_limit, &r->score, &r->te, &r->qe, &bs, &br, &n_b, &cells);
    if (n_b > 0) {
        const int max_sc = o->a > 1 ? o->a : 1;
        const int mx = (r->score + max_sc - 1) / max_sc, low = r->te - mx, high = r->te + mx;
        for (i = 0; i < n_b; ++i)
            if ((br[i] < low || br[i] > high) && bs[i] > r->score2) { r->score2 = bs[i]; r->te2 = br[i]; }
    }
    free(bs); free(br);
    if (minsc < 0 || r->score < minsc || r->te < 0) return cells;
    {
        int sc2, te2, qe2;
        local_pass(r->qe + 1, query + r->qe, r->te + 1, target + r->te, -1, -1, o, no_limit, r->score, &sc2, &te2, &qe2, 0, 0, 0, &cells);
        if (sc2 == r->score) { r->tb = r->te - te2; r->qb = r->qe - qe2; }
    }
    return cells;
}
