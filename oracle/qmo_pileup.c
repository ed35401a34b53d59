/*
 * qmo_pileup.c -- ORACLE (test infrastructure): per-column allele counting with the read admission and
 * mate-overlap rules of `bcftools mpileup -B` (reference call site rules/vcfcall.smk:115; upstream
 * htslib sam.c bam_plp_* / bcftools bam2bcf.c bcf_call_glfgen are not vendored -- SURVEY.md A.8-A.9
 * is the spec).  Depth cap disabled (SURVEY.md A.8 caveat).  PARITY UNPINNED -- see qmo.h.
 */
#include <stdlib.h>
#include <string.h>
#include "qmo.h"

/* mirror of the private layout in qmo_mem.c: only offsets are needed here */
struct qmo_ref { int n_contigs, k; int64_t l_pac, *off, *len; uint8_t *fwd; int64_t n_km; uint64_t *km_key; uint32_t *km_pos; };

void qmo_pileup_opt_default(qmo_pileup_opt_t *p) { p->min_mapq = 0; p->min_bq = 13; p->count_orphans = 0; p->ignore_overlaps = 0; }

#define INC(x) do { _Pragma("omp atomic") (x)++; } while (0)

static int admitted(const qmo_pileup_opt_t *po, const qmo_aln_t *a)
{
    if (a->flag & (0x4 | 0x100 | 0x200 | 0x400)) return 0;
    if (a->n_cigar == 0 || a->n_cigar == 255) return 0;
    if (a->mapq < po->min_mapq) return 0;
    if ((a->flag & 0x1) && !(a->flag & 0x2) && !po->count_orphans) return 0;
    return 1;
}

/* expand an alignment into per-query-base reference positions (-1 = not an M-type base); SEQ order */
static void expand(const qmo_aln_t *a, int l_seq, int32_t *rpos)
{
    int k, x = 0, p = a->pos, i;
    for (i = 0; i < l_seq; ++i) rpos[i] = -1;
    for (k = 0; k < a->n_cigar; ++k) {
        int op = a->cigar[k] & 0xf, len = (int)(a->cigar[k] >> 4);
        if (op == 0) { for (i = 0; i < len; ++i) rpos[x + i] = p + i; x += len; p += len; }
        else if (op == 1 || op == 4) x += len;
        else if (op == 2) p += len;
    }
}

void qmo_pileup(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_pairs, const qmo_aln_t *alns,
                const uint8_t *reads, const uint8_t *quals, int stride, const int32_t *lens, int32_t *counts)
{
    /* pairs are independent and the counts are integer sums: parallel over pairs with atomic increments */
#pragma omp parallel
    {
    int64_t pi;
    int32_t *rp[2];
    uint8_t *sq[2], *ql[2];
    rp[0] = (int32_t *)malloc(4 * stride); rp[1] = (int32_t *)malloc(4 * stride);
    sq[0] = (uint8_t *)malloc(stride); sq[1] = (uint8_t *)malloc(stride);
    ql[0] = (uint8_t *)malloc(stride); ql[1] = (uint8_t *)malloc(stride);
#pragma omp for schedule(dynamic, 1024)
    for (pi = 0; pi < n_pairs; ++pi) {
        const qmo_aln_t *a[2] = { &alns[2 * pi], &alns[2 * pi + 1] };
        int ok[2], e, i, L[2] = { lens[2 * pi], lens[2 * pi + 1] };
        for (e = 0; e < 2; ++e) {
            const uint8_t *rd = reads + (2 * pi + e) * stride, *qv = quals + (2 * pi + e) * stride;
            ok[e] = admitted(po, a[e]);
            if (!ok[e]) continue;
            const int rev = (a[e]->flag & 0x10) != 0;
            for (i = 0; i < L[e]; ++i) {       /* BAM SEQ/QUAL orientation */
                int c = rev ? rd[L[e] - 1 - i] : rd[i];
                sq[e][i] = (uint8_t)(rev ? (c > 3 ? 4 : 3 - c) : c);
                ql[e][i] = rev ? qv[L[e] - 1 - i] : qv[i];
            }
            expand(a[e], L[e], rp[e]);
        }
        /* mate overlap (htslib overlap_push / tweak_overlap_quality): both mates admitted, proper pair,
         * mate mapped, |isize| < 2 l_qseq; `first` is the mate that comes first in coordinate order */
        if (!po->ignore_overlaps && ok[0] && ok[1] && a[0]->rid == a[1]->rid &&
            (a[0]->flag & 0x2) && !(a[0]->flag & 0x8) && abs(a[0]->tlen) < 2 * L[0] && abs(a[1]->tlen) < 2 * L[1]) {
            int first = 0;
            int r0 = (a[0]->flag & 0x10) != 0, r1 = (a[1]->flag & 0x10) != 0;
            if (a[1]->pos < a[0]->pos || (a[1]->pos == a[0]->pos && r1 < r0)) first = 1;
            const int A = first, B = !first;
            int ia, ib = 0;
            for (ia = 0; ia < L[A]; ++ia) {
                int p = rp[A][ia];
                if (p < 0) continue;
                while (ib < L[B] && (rp[B][ib] < 0 || rp[B][ib] < p)) ++ib;
                if (ib >= L[B]) break;
                if (rp[B][ib] != p) continue;
                if (sq[A][ia] == sq[B][ib]) {
                    int q = ql[A][ia] + ql[B][ib];
                    ql[A][ia] = (uint8_t)(q > 200 ? 200 : q); ql[B][ib] = 0;
                } else if (ql[A][ia] >= ql[B][ib]) { ql[A][ia] = (uint8_t)(0.8 * ql[A][ia]); ql[B][ib] = 0; }
                else { ql[B][ib] = (uint8_t)(0.8 * ql[B][ib]); ql[A][ia] = 0; }
            }
        }
        for (e = 0; e < 2; ++e) {
            if (!ok[e]) continue;
            const int rev = (a[e]->flag & 0x10) != 0;
            int32_t *base = counts + (R->off[a[e]->rid]) * QMO_NCH;
            int k, x = 0, p = a[e]->pos, last_m = -1, started = 0;
            for (k = 0; k < a[e]->n_cigar; ++k) {
                int op = a[e]->cigar[k] & 0xf, len = (int)(a[e]->cigar[k] >> 4);
                if (op == 0) {
                    if (!started) { INC(base[(int64_t)p * QMO_NCH + 15]); started = 1; }
                    for (i = 0; i < len; ++i) {
                        int32_t *row = base + (int64_t)(p + i) * QMO_NCH;
                        INC(row[14]);
                        if (ql[e][x + i] >= po->min_bq) INC(row[(rev ? 6 : 0) + sq[e][x + i]]);
                    }
                    x += len; p += len; last_m = p - 1;
                } else if (op == 1) { if (last_m >= 0) INC(base[(int64_t)last_m * QMO_NCH + 12]); x += len; }
                else if (op == 4) x += len;
                else if (op == 2) {
                    if (last_m >= 0) INC(base[(int64_t)last_m * QMO_NCH + 13]);
                    for (i = 0; i < len; ++i) INC(base[(int64_t)(p + i) * QMO_NCH + (rev ? 11 : 5)]);
                    p += len;
                }
            }
        }
    }
    free(rp[0]); free(rp[1]); free(sq[0]); free(sq[1]); free(ql[0]); free(ql[1]);
    }
}
