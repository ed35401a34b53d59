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
