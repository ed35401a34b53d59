/*
 * qmo_pileup.c -- ORACLE (test infrastructure): per-column allele counting with the read admission and
 * mate-overlap rules of `bcftools mpileup -B` (reference call site rules/vcfcall.smk:115; upstream
 * htslib sam.c bam_plp_* / bcftools bam2bcf.c bcf_call_glfgen are not vendored -- SURVEY.md A.8-A.9
 * is the spec).  Depth cap disabled (SURVEY.md A.8 caveat).  PARITY UNPINNED -- see qmo.h.
 */
#include <stdlib.h>
#include <string.h>
#include "qmo_priv.h"

/* mirror of the private layout in qmo_mem.c: only offsets are needed here */

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

/* ---- depth cap: which admitted reads htslib's pileup iterator keeps under `bcftools mpileup -d N`.
 * htslib 1.9 sam.c: bam_plp_push drops a read whose start equals the iterator's current position while more than maxcnt
 * nodes are allocated (the reads buffered and not yet passed, plus the iterator's spare tail node); bam_plp_next frees the
 * reads that ended at or before the position it is assembling, then moves one position on (or jumps to the first buffered
 * read); bam_plp_auto feeds one read whenever the iterator has nothing beyond its position.  Reads come in BAM order
 * (contig, position, strand, input order).  The rule is order dependent by design; this restates it step by step.
 * keep[r] = 1 for reads the iterator keeps, 0 for dropped or not admitted ones.  PARITY UNPINNED (no htslib here). ---- */
typedef struct { int64_t key; int64_t rec; } cap_t;
static int cap_cmp(const void *a, const void *b)
{
    const cap_t *x = (const cap_t *)a, *y = (const cap_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->rec < y->rec ? -1 : x->rec > y->rec;
}
/* min-heap of linked reads on (contig, end): what bam_plp_next releases first */
typedef struct { int tid; int64_t end; int64_t node; } cap_heap_t;
static int heap_less(const cap_heap_t *a, const cap_heap_t *b) { return a->tid != b->tid ? a->tid < b->tid : a->end < b->end; }
static void heap_push(cap_heap_t *h, int64_t *n, cap_heap_t v)
{
    int64_t i = (*n)++;
    while (i > 0 && heap_less(&v, &h[(i - 1) >> 1])) { h[i] = h[(i - 1) >> 1]; i = (i - 1) >> 1; }
    h[i] = v;
}
static void heap_pop(cap_heap_t *h, int64_t *n)
{
    const cap_heap_t v = h[--(*n)];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= *n) break;
        if (c + 1 < *n && heap_less(&h[c + 1], &h[c])) ++c;
        if (!heap_less(&h[c], &v)) break;
        h[i] = h[c]; i = c;
    }
    h[i] = v;
}
void qmo_depth_cap(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_reads, const qmo_aln_t *alns, int max_depth, uint8_t *keep)
{
    cap_t *ord = (cap_t *)malloc(sizeof(cap_t) * (size_t)(n_reads ? n_reads : 1));
    int64_t m = 0, r, i;
    (void)R;
    memset(keep, 0, (size_t)n_reads);
    for (r = 0; r < n_reads; ++r)
        if (admitted(po, &alns[r])) {
            ord[m].key = ((int64_t)alns[r].rid << 40) | ((int64_t)alns[r].pos << 1) | ((alns[r].flag & 0x10) != 0);
            ord[m].rec = r; ++m;
        }
    qsort(ord, (size_t)m, sizeof(cap_t), cap_cmp);
    {
        int it_tid = 0, max_tid = -1;
        int64_t it_pos = 0, max_pos = -1, cnt = 1 /* the iterator's spare tail node */;
        cap_heap_t *heap = (cap_heap_t *)malloc(sizeof(cap_heap_t) * (size_t)(m ? m : 1));
        int64_t n_heap = 0, head = 0, n_link = 0;
        int64_t *lbeg = (int64_t *)malloc(8 * (size_t)(m ? m : 1));
        int *ltid = (int *)malloc(4 * (size_t)(m ? m : 1));
        uint8_t *gone = (uint8_t *)calloc((size_t)(m ? m : 1), 1);
        for (i = 0; i < m; ++i) {
            const qmo_aln_t *a = &alns[ord[i].rec];
            int k; int64_t rlen = 0, end;
            for (k = 0; k < a->n_cigar; ++k) { const int op = a->cigar[k] & 0xf; if (op == 0 || op == 2) rlen += (int64_t)(a->cigar[k] >> 4); }
            end = a->pos + (rlen > 0 ? rlen : 1);
            /* bam_plp_next: as long as something beyond the iterator's position is buffered */
            while (max_tid > it_tid || (max_tid == it_tid && max_pos > it_pos)) {
                while (n_heap && (heap[0].tid < it_tid || (heap[0].tid == it_tid && heap[0].end <= it_pos))) {      /* release */
                    gone[heap[0].node] = 1; --cnt;
                    heap_pop(heap, &n_heap);
                }
                while (head < n_link && gone[head]) ++head;
                if (head < n_link && it_tid < ltid[head]) { it_tid = ltid[head]; it_pos = lbeg[head]; }          /* next contig */
                else if (head < n_link && it_pos < lbeg[head]) it_pos = lbeg[head];                              /* jump a gap */
                else ++it_pos;
            }
            /* bam_plp_push */
            if (it_tid == a->rid && it_pos == a->pos && cnt > max_depth) continue;                 /* dropped */
            keep[ord[i].rec] = 1;
            max_tid = a->rid; max_pos = a->pos;
            if (end > it_pos || a->rid > it_tid) {                                                  /* linked: a node is taken */
                cap_heap_t v;
                v.tid = a->rid; v.end = end; v.node = n_link;
                heap_push(heap, &n_heap, v);
                lbeg[n_link] = a->pos; ltid[n_link] = a->rid; ++n_link; ++cnt;
            }
        }
        free(heap); free(lbeg); free(ltid); free(gone);
    }
    free(ord);
}

/* ---- indel alleles ("indel alleles into a small hash table"): every insertion / deletion operation of an
 * admitted read's CIGAR, keyed by (anchor = the reference base in front of the event, type, length clamped to 255, the first 11
 * inserted bases as the forward strand reads them + an "N among them" flag), forward and reverse reads counted apart.  Same
 * admission as the counting above.  Records come out sorted by key.  out: 10 int32 per record {rid, pos, len, type, has_n, seq,
 * n_fwd, n_rev, key_lo, key_hi}.  Returns the number of alleles (<= max_out are written). ---- */
typedef struct { uint64_t key; int32_t rev; } ind_t;
static int ind_cmp(const void *a, const void *b)
{
    const ind_t *x = (const ind_t *)a, *y = (const ind_t *)b;
    return x->key < y->key ? -1 : x->key > y->key;
}
int64_t qmo_indels(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_pairs, const qmo_aln_t *alns,
                   const uint8_t *reads, int stride, const int32_t *lens, int32_t *out, int64_t max_out)
{
    int64_t r, n = 0, cap = 1024, m = 0, i;
    ind_t *ev = (ind_t *)malloc(sizeof(ind_t) * (size_t)cap);
    for (r = 0; r < 2 * n_pairs; ++r) {
        const qmo_aln_t *a = &alns[r];
        const uint8_t *rd = reads + r * stride;
        const int L = lens[r], rev = (a->flag & 0x10) != 0;
        int k, x = 0, p = a->pos, last_m = -1;
        if (!admitted(po, a)) continue;
        for (k = 0; k < a->n_cigar; ++k) {
            const int op = a->cigar[k] & 0xf, len = (int)(a->cigar[k] >> 4);
            if (op == 0) { x += len; p += len; last_m = p - 1; }
            else if (op == 4) x += len;
            else if (op == 1 || op == 2) {
                if (last_m >= 0) {
                    uint32_t seq = 0; int has_n = 0, j;
                    if (op == 1)
                        for (j = 0; j < len && j < 11; ++j) {
                            int c = rev ? rd[L - 1 - (x + j)] : rd[x + j];
                            c = rev ? (c > 3 ? 4 : 3 - c) : c;
                            if (c > 3) has_n = 1; else seq |= (uint32_t)c << (2 * j);
                        }
                    if (n == cap) { cap *= 2; ev = (ind_t *)realloc(ev, sizeof(ind_t) * (size_t)cap); }
                    ev[n].key = (uint64_t)(uint32_t)(R->off[a->rid] + last_m) << 32 | (uint64_t)(op == 2) << 31 |
                                (uint64_t)(len > 255 ? 255 : len) << 23 | (uint64_t)has_n << 22 | (uint64_t)(seq & 0x3fffffu);
                    ev[n].rev = rev;
                    ++n;
                }
                if (op == 1) x += len; else p += len;
            }
        }
    }
    qsort(ev, (size_t)n, sizeof(ind_t), ind_cmp);
    for (i = 0; i < n;) {
        int64_t j = i;
        int32_t nf = 0, nr = 0;
        while (j < n && ev[j].key == ev[i].key) { if (ev[j].rev) ++nr; else ++nf; ++j; }
        if (m < max_out) {
            const uint64_t key = ev[i].key;
            const int64_t g = (int64_t)(key >> 32);
            int rid = 0, c;
            int32_t *o = out + 10 * m;
            for (c = 0; c < R->n_contigs; ++c) if (g >= R->off[c] && g < R->off[c] + R->len[c]) rid = c;
            o[0] = rid; o[1] = (int32_t)(g - R->off[rid]); o[2] = (int32_t)((key >> 23) & 0xff); o[3] = (int32_t)((key >> 31) & 1);
            o[4] = (int32_t)((key >> 22) & 1); o[5] = (int32_t)(key & 0x3fffffu); o[6] = nf; o[7] = nr;
            o[8] = (int32_t)(uint32_t)key; o[9] = (int32_t)(uint32_t)(key >> 32);
        }
        ++m;
        i = j;
    }
    free(ev);
    return m;
}

/* ---- samtools mpileup text (reference call site rules/vcfcall.smk:39, consumer VarScan; upstream samtools 1.9
 * bam_plcmd.c mpileup / pileup_seq and htslib sam.c resolve_cigar2 are not vendored -- SURVEY.md A.10 is the spec).
 * Same read admission and mate-overlap quality rewrite as the counting above (BAQ off, no depth cap).  Reads enter a
 * column in coordinate-sorted order (samtools sort: contig, position, strand; input order on ties).  A line is
 * written for every column some admitted read covers: chrom, 1-based position, reference base, the number of
 * entries whose base quality reaches min_bq, their base string and their quality string.  Per entry:
 *   ^X   the read starts here (X = min(MAPQ, 93) + 33)
 *   . ,  base equal to the reference (forward / reverse strand);  ACGTN acgtn otherwise;  * deleted here
 *   +nSEQ / -nSEQ  an insertion / deletion follows this position (inserted read bases / deleted reference bases)
 *   $    the read ends here
 * The quality of a deleted position is the one of the first read base behind the deletion (resolve_cigar2 leaves qpos
 * there).  PARITY UNPINNED. */
typedef struct { char *s; int n, cap; } dstr_t;
static void dput(dstr_t *d, int c)
{
    if (d->n == d->cap) { d->cap = d->cap ? d->cap * 2 : 16; d->s = (char *)realloc(d->s, (size_t)d->cap); }
    d->s[d->n++] = (char)c;
}
static void dnum(dstr_t *d, int v) { char b[16]; int n = 0; do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v); while (n) dput(d, b[--n]); }

typedef struct { int64_t key; int32_t rec; } srt_t;
static int srt_cmp(const void *a, const void *b)
{
    const srt_t *x = (const srt_t *)a, *y = (const srt_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->rec < y->rec ? -1 : x->rec > y->rec;
}

int64_t qmo_mpileup_text(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_pairs, const qmo_aln_t *alns,
                         const uint8_t *reads, const uint8_t *quals, int stride, const int32_t *lens,
                         const char *const *names, char **out)
{
    const int64_t n = 2 * n_pairs;
    uint8_t *sq = (uint8_t *)malloc((size_t)n * stride), *ql = (uint8_t *)malloc((size_t)n * stride);
    uint8_t *ok = (uint8_t *)calloc((size_t)n, 1);
    int32_t *rpa = (int32_t *)malloc(4 * (size_t)stride), *rpb = (int32_t *)malloc(4 * (size_t)stride);
    dstr_t *bs = (dstr_t *)calloc((size_t)R->l_pac, sizeof(dstr_t)), *qs = (dstr_t *)calloc((size_t)R->l_pac, sizeof(dstr_t));
    uint8_t *covered = (uint8_t *)calloc((size_t)R->l_pac, 1);
    srt_t *order = (srt_t *)malloc(sizeof(srt_t) * (size_t)n);
    int64_t pi, r, m = 0, c, total = 0;
    static const char up[] = "ACGTN", lo[] = "acgtn";
    for (pi = 0; pi < n_pairs; ++pi) {
        const qmo_aln_t *a[2] = { &alns[2 * pi], &alns[2 * pi + 1] };
        int e, i, L[2] = { lens[2 * pi], lens[2 * pi + 1] };
        for (e = 0; e < 2; ++e) {
            const uint8_t *rd = reads + (2 * pi + e) * stride, *qv = quals + (2 * pi + e) * stride;
            uint8_t *s = sq + (2 * pi + e) * stride, *q = ql + (2 * pi + e) * stride;
            const int rev = (a[e]->flag & 0x10) != 0;
            ok[2 * pi + e] = (uint8_t)admitted(po, a[e]);
            for (i = 0; i < L[e]; ++i) {
                int cc = rev ? rd[L[e] - 1 - i] : rd[i];
                s[i] = (uint8_t)(rev ? (cc > 3 ? 4 : 3 - cc) : cc);
                q[i] = rev ? qv[L[e] - 1 - i] : qv[i];
            }
        }
        if (!po->ignore_overlaps && ok[2 * pi] && ok[2 * pi + 1] && a[0]->rid == a[1]->rid &&
            (a[0]->flag & 0x2) && !(a[0]->flag & 0x8) && abs(a[0]->tlen) < 2 * L[0] && abs(a[1]->tlen) < 2 * L[1]) {
            int first = 0, r0 = (a[0]->flag & 0x10) != 0, r1 = (a[1]->flag & 0x10) != 0, ia, ib = 0;
            if (a[1]->pos < a[0]->pos || (a[1]->pos == a[0]->pos && r1 < r0)) first = 1;
            {
                const int A = first, B = !first;
                uint8_t *sA = sq + (2 * pi + A) * stride, *sB = sq + (2 * pi + B) * stride;
                uint8_t *qA = ql + (2 * pi + A) * stride, *qB = ql + (2 * pi + B) * stride;
                expand(a[A], L[A], rpa); expand(a[B], L[B], rpb);
                for (ia = 0; ia < L[A]; ++ia) {
                    int p = rpa[ia];
                    if (p < 0) continue;
                    while (ib < L[B] && (rpb[ib] < 0 || rpb[ib] < p)) ++ib;
                    if (ib >= L[B]) break;
                    if (rpb[ib] != p) continue;
                    if (sA[ia] == sB[ib]) { int q = qA[ia] + qB[ib]; qA[ia] = (uint8_t)(q > 200 ? 200 : q); qB[ib] = 0; }
                    else if (qA[ia] >= qB[ib]) { qA[ia] = (uint8_t)(0.8 * qA[ia]); qB[ib] = 0; }
                    else { qB[ib] = (uint8_t)(0.8 * qB[ib]); qA[ia] = 0; }
                }
            }
        }
    }
    for (r = 0; r < n; ++r)
        if (ok[r]) { order[m].key = ((R->off[alns[r].rid] + alns[r].pos) << 1) | ((alns[r].flag & 0x10) != 0); order[m].rec = (int32_t)r; ++m; }
    qsort(order, (size_t)m, sizeof(srt_t), srt_cmp);
    for (r = 0; r < m; ++r) {
        const qmo_aln_t *a = &alns[order[r].rec];
        const uint8_t *s = sq + (int64_t)order[r].rec * stride, *q = ql + (int64_t)order[r].rec * stride;
        const int L = lens[order[r].rec], rev = (a->flag & 0x10) != 0;
        const int64_t base = R->off[a->rid], clen = R->len[a->rid];
        const char *let = rev ? lo : up;
        int k, x = a->pos, y = 0, end = a->pos, i, j;
        for (k = 0; k < a->n_cigar; ++k) { int op = a->cigar[k] & 0xf; if (op == 0 || op == 2) end += (int)(a->cigar[k] >> 4); }
        for (k = 0; k < a->n_cigar; ++k) {
            const int op = a->cigar[k] & 0xf, len = (int)(a->cigar[k] >> 4);
            if (op == 1 || op == 4) { y += len; continue; }
            if (op != 0 && op != 2) continue;
            for (i = 0; i < len; ++i) {
                const int p = x + i, is_del = op == 2, qpos = is_del ? y : y + i;
                const int qual = qpos < L ? q[qpos] : 0;
                int indel = 0;
                covered[base + p] = 1;
                if (i == len - 1 && k + 1 < a->n_cigar) {
                    const int op2 = a->cigar[k + 1] & 0xf, l2 = (int)(a->cigar[k + 1] >> 4);
                    if (op2 == 2) indel = -l2; else if (op2 == 1) indel = l2;
                }
                if (qual < po->min_bq) continue;
                {
                    dstr_t *d = &bs[base + p];
                    if (p == a->pos) { dput(d, '^'); dput(d, a->mapq > 93 ? 126 : a->mapq + 33); }
                    if (!is_del) {
                        const int cc = qpos < L ? s[qpos] : 4;
                        if (cc < 4 && cc == R->fwd[base + p]) dput(d, rev ? ',' : '.'); else dput(d, let[cc]);
                    } else dput(d, '*');
                    if (indel > 0) { dput(d, '+'); dnum(d, indel); for (j = 1; j <= indel; ++j) dput(d, let[qpos + j < L ? s[qpos + j] : 4]); }
                    else if (indel < 0) { dput(d, '-'); dnum(d, -indel); for (j = 1; j <= -indel; ++j) dput(d, p + j < clen ? let[R->fwd[base + p + j]] : let[4]); }
                    if (p == end - 1) dput(d, '$');
                    dput(&qs[base + p], qual + 33 < 126 ? qual + 33 : 126);
                }
            }
            x += len; if (op == 0) y += len;
        }
    }
    {
        dstr_t o = {0, 0, 0};
        int ci;
        for (ci = 0; ci < R->n_contigs; ++ci)
            for (c = 0; c < R->len[ci]; ++c) {
                const int64_t g = R->off[ci] + c;
                const char *nm = names[ci];
                int i;
                if (!covered[g]) continue;
                while (*nm) dput(&o, *nm++);
                dput(&o, '\t'); dnum(&o, (int)(c + 1)); dput(&o, '\t'); dput(&o, up[R->fwd[g]]); dput(&o, '\t'); dnum(&o, qs[g].n); dput(&o, '\t');
                /* samtools 1.9 bam_plcmd.c: a column whose reads all failed -Q prints "*" for both strings */
                for (i = 0; i < bs[g].n; ++i) dput(&o, bs[g].s[i]);
                if (qs[g].n == 0) dput(&o, '*');
                dput(&o, '\t');
                for (i = 0; i < qs[g].n; ++i) dput(&o, qs[g].s[i]);
                if (qs[g].n == 0) dput(&o, '*');
                dput(&o, '\n');
            }
        *out = o.s; total = o.n;
    }
    for (c = 0; c < R->l_pac; ++c) { free(bs[c].s); free(qs[c].s); }
    free(bs); free(qs); free(covered); free(order); free(sq); free(ql); free(ok); free(rpa); free(rpb);
    return total;
}
void qmo_free(void *p) { free(p); }
