/*
 * qmo_baq.c -- ORACLE (test infrastructure): base alignment quality, the per-base quality cap that `bcftools mpileup` and
 * `samtools mpileup` apply to every read unless -B is given (reference call sites rules/vcfcall.smk:39 and :115 -- the
 * reference passes no -B, so its runs have BAQ ON).
 *
 * Upstream is htslib 1.9 (conda pin config/conda_env.yaml): realn.c sam_prob_realn(b, ref, ref_len, flag = 3: apply +
 * extended BAQ, what mplp_func passes) and probaln.c kpa_glocal (banded profile HMM, forward / backward / posterior
 * decoding, double precision, per-row rescaling).  htslib is neither vendored under /root/reference nor installed, and the
 * reference holds no BAQ test vector: this file restates the published algorithm statement by statement (operation order
 * kept, because the result is rounded to integers from double-precision sums) -- PARITY UNPINNED, like the rest of the
 * pileup oracle (see qmo.h).
 *
 * What the two functions do:
 *   qmo_kpa_glocal   HMM with states M / I / D per (query base i, reference base k) inside a band of half-width bw around
 *                    the diagonal; emission 1 - err / err/3 for match / mismatch with err = 10^(-Q/10) (stored as float, as
 *                    upstream does), 1/4 for insertions; gap open d = 0.001, extension e = 0.1.  Returns for every query base
 *                    the most probable state (reference offset << 2 | 0 = M, 1 = I) and its posterior as a phred value.
 *   qmo_baq          per admitted read: the reference window around the alignment, kpa_glocal, then for every aligned (M) base
 *                    BAQ = posterior phred if the HMM agrees with the CIGAR's placement of the base, else 0; "extended" BAQ
 *                    takes, inside each M block, min(running max from the left, running max from the right); the base
 *                    quality becomes min(quality, BAQ).  Bases outside M blocks keep their quality.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "qmo_priv.h"

#define KPA_EI .25
#define KPA_EM .33333333333

#define set_u(u, b, i, k) { int x_ = (i) - (b); x_ = x_ > 0 ? x_ : 0; (u) = ((k) - x_ + 1) * 3; }

static double g_qual2prob[256];

/* ref / query: base codes 0..3, anything above = ambiguous.  state / q: l_query entries.  Returns the phred-scaled
 * likelihood (upstream's return value; unused by BAQ). */
int qmo_kpa_glocal(const uint8_t *_ref, int l_ref, const uint8_t *_query, int l_query, const uint8_t *iqual,
                   double cd, double ce, int cbw, int *state, uint8_t *q)
{
    double **f, **b, *s, m[9], sI, sM, bI, bM;
    float *qual, *_qual;
    const uint8_t *ref, *query;
    int bw, bw2, i, k, Pr;

    if (l_ref <= 0 || l_query <= 0) return 0;
    ref = _ref - 1; query = _query - 1;                 /* 1-based coordinates */
    bw = l_ref > l_query ? l_ref : l_query;
    if (bw > cbw) bw = cbw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    bw2 = bw * 2 + 1;
    f = (double **)calloc(l_query + 1, sizeof(double *));
    b = (double **)calloc(l_query + 1, sizeof(double *));
    for (i = 0; i <= l_query; ++i) {
        f[i] = (double *)calloc(bw2 * 3 + 6, sizeof(double));
        b[i] = (double *)calloc(bw2 * 3 + 6, sizeof(double));
    }
    s = (double *)calloc(l_query + 2, sizeof(double));  /* the rows' scaling factors */
    _qual = (float *)calloc(l_query, sizeof(float));
    if (g_qual2prob[0] == 0)
        for (i = 0; i < 256; ++i) g_qual2prob[i] = pow(10, -i / 10.);
    for (i = 0; i < l_query; ++i) _qual[i] = (float)g_qual2prob[iqual ? iqual[i] : 30];
    qual = _qual - 1;
    /* transition probabilities */
    sM = sI = 1. / (2 * l_query + 2);
    m[0 * 3 + 0] = (1 - cd - cd) * (1 - sM); m[0 * 3 + 1] = m[0 * 3 + 2] = cd * (1 - sM);
    m[1 * 3 + 0] = (1 - ce) * (1 - sI); m[1 * 3 + 1] = ce * (1 - sI); m[1 * 3 + 2] = 0.;
    m[2 * 3 + 0] = 1 - ce; m[2 * 3 + 1] = 0.; m[2 * 3 + 2] = ce;
    bM = (1 - cd) / l_ref; bI = cd / l_ref;
    /*** forward ***/
    set_u(k, bw, 0, 0);
    f[0][k] = s[0] = 1.;
    {   /* f[1] */
        double *fi = f[1], sum;
        int beg = 1, end = l_ref < bw + 1 ? l_ref : bw + 1, _beg, _end;
        for (k = beg, sum = 0.; k <= end; ++k) {
            int u;
            double e = (ref[k] > 3 || query[1] > 3) ? 1. : ref[k] == query[1] ? 1. - qual[1] : qual[1] * KPA_EM;
            set_u(u, bw, 1, k);
            fi[u + 0] = e * bM; fi[u + 1] = KPA_EI * bI;
            sum += fi[u] + fi[u + 1];
        }
        s[1] = sum;
        set_u(_beg, bw, 1, beg); set_u(_end, bw, 1, end); _end += 2;
        for (k = _beg; k <= _end; ++k) fi[k] /= sum;
    }
    for (i = 2; i <= l_query; ++i) {   /* f[2 .. l_query] */
        double *fi = f[i], *fi1 = f[i - 1], sum, qli = qual[i];
        int beg = 1, end = l_ref, x, _beg, _end;
        uint8_t qyi = query[i];
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg, sum = 0.; k <= end; ++k) {
            int u, v11, v01, v10;
            double e;
            e = (ref[k] > 3 || qyi > 3) ? 1. : ref[k] == qyi ? 1. - qli : qli * KPA_EM;
            set_u(u, bw, i, k); set_u(v11, bw, i - 1, k - 1); set_u(v10, bw, i - 1, k); set_u(v01, bw, i, k - 1);
            fi[u + 0] = e * (m[0] * fi1[v11 + 0] + m[3] * fi1[v11 + 1] + m[6] * fi1[v11 + 2]);
            fi[u + 1] = KPA_EI * (m[1] * fi1[v10 + 0] + m[4] * fi1[v10 + 1]);
            fi[u + 2] = m[2] * fi[v01 + 0] + m[8] * fi[v01 + 2];
            sum += fi[u] + fi[u + 1] + fi[u + 2];
        }
        s[i] = sum;
        set_u(_beg, bw, i, beg); set_u(_end, bw, i, end); _end += 2;
        for (k = _beg, sum = 1. / sum; k <= _end; ++k) fi[k] *= sum;
    }
    {   /* f[l_query + 1] */
        double sum;
        for (k = 1, sum = 0.; k <= l_ref; ++k) {
            int u;
            set_u(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += f[l_query][u + 0] * sM + f[l_query][u + 1] * sI;
        }
        s[l_query + 1] = sum;
    }
    {   /* likelihood */
        double p = 1., Pr1 = 0.;
        for (i = 0; i <= l_query + 1; ++i) {
            p *= s[i];
            if (p < 1e-100) Pr1 += -4.343 * log(p), p = 1.;
        }
        Pr1 += -4.343 * log(p * l_ref * l_query);
        Pr = (int)(Pr1 + .499);
    }
    /*** backward ***/
    for (k = 1; k <= l_ref; ++k) {   /* b[l_query] */
        int u;
        double *bi = b[l_query];
        set_u(u, bw, l_query, k);
        if (u < 3 || u >= bw2 * 3 + 3) continue;
        bi[u + 0] = sM / s[l_query] / s[l_query + 1]; bi[u + 1] = sI / s[l_query] / s[l_query + 1];
    }
    for (i = l_query - 1; i >= 1; --i) {   /* b[l_query - 1 .. 1] */
        int beg = 1, end = l_ref, x, _beg, _end;
        double *bi = b[i], *bi1 = b[i + 1], y = (i > 1), qli1 = qual[i + 1];
        uint8_t qyi1 = query[i + 1];
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = end; k >= beg; --k) {
            int u, v11, v01, v10;
            double e;
            set_u(u, bw, i, k); set_u(v11, bw, i + 1, k + 1); set_u(v10, bw, i + 1, k); set_u(v01, bw, i, k + 1);
            e = (k >= l_ref ? 0 : (ref[k + 1] > 3 || qyi1 > 3) ? 1. : ref[k + 1] == qyi1 ? 1. - qli1 : qli1 * KPA_EM) * bi1[v11];
            bi[u + 0] = e * m[0] + KPA_EI * m[1] * bi1[v10 + 1] + m[2] * bi[v01 + 2];
            bi[u + 1] = e * m[3] + KPA_EI * m[4] * bi1[v10 + 1];
            bi[u + 2] = (e * m[6] + m[8] * bi[v01 + 2]) * y;
        }
        set_u(_beg, bw, i, beg); set_u(_end, bw, i, end); _end += 2;
        for (k = _beg, y = 1. / s[i]; k <= _end; ++k) bi[k] *= y;
    }
    /* (upstream also computes b[0] as a self-check that is not part of the result) */
    /*** posterior decoding ***/
    for (i = 1; i <= l_query; ++i) {
        double sum = 0., *fi = f[i], *bi = b[i], max = 0.;
        int beg = 1, end = l_ref, x, max_k = -1;
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg; k <= end; ++k) {
            int u;
            double z;
            set_u(u, bw, i, k);
            z = fi[u + 0] * bi[u + 0]; if (z > max) max = z, max_k = (k - 1) << 2 | 0; sum += z;
            z = fi[u + 1] * bi[u + 1]; if (z > max) max = z, max_k = (k - 1) << 2 | 1; sum += z;
        }
        max /= sum;
        if (state) state[i - 1] = max_k;
        if (q) { k = (int)(-4.343 * log(1. - max) + .499); q[i - 1] = (uint8_t)(k > 100 ? 99 : k); }
    }
    for (i = 0; i <= l_query; ++i) { free(f[i]); free(b[i]); }
    free(f); free(b); free(s); free(_qual);
    return Pr;
}

static int baq_admitted(const qmo_pileup_opt_t *po, const qmo_aln_t *a)
{
    /* the reads the pileup counts (qmo_pileup.c admitted()); BAQ of any other read is never looked at */
    if (a->flag & (0x4 | 0x100 | 0x200 | 0x400)) return 0;
    if (a->n_cigar == 0 || a->n_cigar == 255) return 0;
    if (a->mapq < po->min_mapq) return 0;
    if ((a->flag & 0x1) && !(a->flag & 0x2) && !po->count_orphans) return 0;
    return 1;
}

/* BAQ of one read in BAM orientation: seq / qual are the record's SEQ / QUAL (reference strand), ref = the contig, codes 0..4.
 * qual is rewritten in place (sam_prob_realn with apply_baq; extend = the flag's bit 1). */
void qmo_baq_read(const uint8_t *ref, int64_t ref_len, const qmo_aln_t *a, int l_qseq, const uint8_t *seq, uint8_t *qual, int extend)
{
    int k, i, bw, x, y, yb, ye, xb, xe;
    if (l_qseq == 0) return;
    x = a->pos; y = 0; yb = ye = xb = xe = -1;
    for (k = 0; k < a->n_cigar; ++k) {
        int op = a->cigar[k] & 0xf, l = (int)(a->cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (yb < 0) yb = y;
            if (xb < 0) xb = x;
            ye = y + l; xe = x + l;
            x += l; y += l;
        } else if (op == 4 || op == 1) y += l;
        else if (op == 2) x += l;
        else if (op == 3) return;
    }
    if (xb < 0) return;                                   /* no aligned base (cannot happen for a mapped record) */
    bw = 7;
    if (abs((xe - xb) - (ye - yb)) > bw) bw = abs((xe - xb) - (ye - yb)) + 3;
    xb -= yb + bw / 2; if (xb < 0) xb = 0;
    xe += l_qseq - ye + bw / 2;
    if (xe - xb - l_qseq > bw) {                          /* (the second statement sees the first one's xb, as upstream) */
        xb += (xe - xb - l_qseq - bw) / 2;
        xe -= (xe - xb - l_qseq - bw) / 2;
    }
    {
        uint8_t *bq = (uint8_t *)calloc(l_qseq + 1, 1), *q = (uint8_t *)calloc(l_qseq, 1), *r;
        int *state = (int *)calloc(l_qseq, sizeof(int));
        memcpy(bq, qual, l_qseq);
        if (xe > ref_len) xe = (int)ref_len;
        if (xe - xb <= 0) { free(bq); free(q); free(state); return; }      /* (a window past the contig: cannot happen for a placed read) */
        r = (uint8_t *)calloc(xe - xb, 1);
        for (i = xb; i < xe; ++i) r[i - xb] = ref[i] > 3 ? 4 : ref[i];
        qmo_kpa_glocal(r, xe - xb, seq, l_qseq, qual, 0.001, 0.1, bw, state, q);
        if (!extend) {
            for (k = 0, x = a->pos, y = 0; k < a->n_cigar; ++k) {
                int op = a->cigar[k] & 0xf, l = (int)(a->cigar[k] >> 4);
                if (op == 0 || op == 7 || op == 8) {
                    for (i = y; i < y + l; ++i) {
                        if ((state[i] & 3) != 0 || state[i] >> 2 != x - xb + (i - y)) bq[i] = 0;
                        else bq[i] = bq[i] < q[i] ? bq[i] : q[i];
                    }
                    x += l; y += l;
                } else if (op == 4 || op == 1) y += l;
                else if (op == 2) x += l;
            }
            for (i = 0; i < l_qseq; ++i) bq[i] = (uint8_t)(qual[i] - bq[i] + 64);
        } else {
            uint8_t *left = (uint8_t *)calloc(l_qseq, 1), *rght = (uint8_t *)calloc(l_qseq, 1);
            for (k = 0, x = a->pos, y = 0; k < a->n_cigar; ++k) {
                int op = a->cigar[k] & 0xf, l = (int)(a->cigar[k] >> 4);
                if (op == 0 || op == 7 || op == 8) {
                    for (i = y; i < y + l; ++i)
                        bq[i] = ((state[i] & 3) != 0 || state[i] >> 2 != x - xb + (i - y)) ? 0 : q[i];
                    for (left[y] = bq[y], i = y + 1; i < y + l; ++i) left[i] = bq[i] > left[i - 1] ? bq[i] : left[i - 1];
                    for (rght[y + l - 1] = bq[y + l - 1], i = y + l - 2; i >= y; --i) rght[i] = bq[i] > rght[i + 1] ? bq[i] : rght[i + 1];
                    for (i = y; i < y + l; ++i) bq[i] = left[i] < rght[i] ? left[i] : rght[i];
                    x += l; y += l;
                } else if (op == 4 || op == 1) y += l;
                else if (op == 2) x += l;
            }
            for (i = 0; i < l_qseq; ++i) bq[i] = (uint8_t)(64 + (qual[i] <= bq[i] ? 0 : qual[i] - bq[i]));
            free(left); free(rght);
        }
        for (i = 0; i < l_qseq; ++i) qual[i] = (uint8_t)(qual[i] - (bq[i] - 64));
        free(bq); free(q); free(r); free(state);
    }
}

/* BAQ of a batch: quals_out = quals with every ADMITTED read's qualities capped (reads as sequenced, like quals).
 * flag: 3 = what bcftools / samtools mpileup pass (apply + extended), 1 = plain BAQ. */
void qmo_baq(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_reads, const qmo_aln_t *alns, const uint8_t *reads,
             const uint8_t *quals, int stride, const int32_t *lens, int flag, uint8_t *quals_out)
{
    int64_t r;
    memcpy(quals_out, quals, (size_t)n_reads * stride);
#pragma omp parallel for schedule(dynamic, 256)
    for (r = 0; r < n_reads; ++r) {
        const qmo_aln_t *a = &alns[r];
        const int L = lens[r];
        if (!baq_admitted(po, a) || L <= 0) continue;
        const int rev = (a->flag & 0x10) != 0;
        uint8_t *sq = (uint8_t *)malloc(L), *ql = (uint8_t *)malloc(L);
        const uint8_t *rd = reads + r * stride, *qv = quals + r * stride;
        int i;
        for (i = 0; i < L; ++i) {                         /* BAM SEQ / QUAL orientation */
            int c = rev ? rd[L - 1 - i] : rd[i];
            sq[i] = (uint8_t)(rev ? (c > 3 ? 4 : 3 - c) : (c > 3 ? 4 : c));
            ql[i] = rev ? qv[L - 1 - i] : qv[i];
        }
        qmo_baq_read(R->fwd + R->off[a->rid], R->len[a->rid], a, L, sq, ql, (flag >> 1) & 1);
        for (i = 0; i < L; ++i) quals_out[r * stride + (rev ? L - 1 - i : i)] = ql[i];
        free(sq); free(ql);
    }
}
