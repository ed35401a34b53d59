/*
 * qmo.h -- CPU ORACLE for the QuasiModo read-level hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain-C restatement of the algorithms the reference pipeline runs through
 * un-vendored third-party binaries (bwa 0.7.17, samtools/bcftools 1.9; call sites
 * rules/bwa.smk:15-18, rules/vcfcall.smk:39,115-117) plus the two in-tree scripts
 * program/extract_TP_FP_SNPs.py:12-57 and scripts/caller_performance_compare.R:29-55,77-143.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (quasimodo_b200/) never links or calls it.
 *
 * PARITY STATUS
 *   - alignment (ksw_extend2 / ksw_global2 / bwa-mem glue), pileup counting: PARITY UNPINNED.
 *     bwa / samtools / bcftools are absent from the image and from /root/reference; the reference
 *     holds no test, fixture or golden vector for them.  The restatement
 *     follows the published algorithms and is held against second, separately
 *     written restatements, stage by stage (tests/test_oracle_ksw.py, tests/test_oracle_mem.py):
 *     oracle/ksw_py.py (ksw_extend2, ksw_align2, ksw_global2), pileup_py.py (counting, indel tally,
 *     text pileup), pestat_py.py (mem_pestat), mapq_py.py + pair_py.py (primary marking, MAPQ,
 *     mem_pair and mem_sam_pe's decision), cigar_py.py (mem_reg2aln / bwa_gen_cigar2), rescue_py.py (mem_matesw, mem_sort_dedup_patch),
 *     depthcap_py.py (htslib's depth cap), a brute-force seed
 *     enumeration; FM-index construction is PINNED to bwa's own ref/X.bwt and ref/X.sa bytes.
 *   - evaluation (TP/FP split, caller_performance table): PINNED against the reference's own
 *     program/extract_TP_FP_SNPs.py run in the build container (tests/golden/eval_*).
 *   - truth VCF: produced by the reference's own program/mummer2vcf.py (tests/golden/make_truth.py).
 */
#ifndef QMO_H
#define QMO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- scoring / bwa-mem options (bwa mem -k 31 defaults) ---- */
typedef struct {
    int32_t a, b;                 /* match score, mismatch penalty                      */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t w;                    /* band width                                         */
    int32_t zdrop;
    int32_t pen_clip5, pen_clip3; /* end bonuses                                        */
    int32_t min_seed_len;         /* -k                                                 */
    int32_t max_occ;              /* k-mers with more occurrences are ignored           */
    int32_t T;                    /* minimum score to output                            */
    int32_t pen_unpaired;
    int32_t max_ins;
    int32_t max_chain_gap;
    int32_t mapq_coef_len;
    float   mask_level, drop_ratio, mask_level_redun;
    int32_t min_chain_weight;
    int32_t flags;                /* QMO_F_*                                            */
    int32_t reserved[2];
} qmo_opt_t;
#define QMO_F_NO_RESCUE 1         /* bwa mem -S: skip mate rescue                       */
#define QMO_F_PATCH     4         /* mem_patch_reg: two collinear hits merged when one global alignment explains them */
#define QMO_F_FM_SEEDS  2         /* seeds from bwa's FM-index (qmo_ref_set_fm) instead of the k-mer hash index */

void qmo_opt_default(qmo_opt_t *o);

/* ---- ksw ---- */
typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } qmo_ext_t;

/* returns executed inner-loop cells (sum over rows of end-beg) */
int64_t qmo_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                        const qmo_opt_t *o, int w, int end_bonus, int h0, qmo_ext_t *out);

/* banded global alignment with traceback; cigar ops as len<<4|op (M=0,I=1,D=2).
 * returns score; *n_cigar <= max_cigar or -1 on overflow */
int qmo_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                    const qmo_opt_t *o, int w, int *n_cigar, uint32_t *cigar, int max_cigar);

/* local alignment of a whole query inside a target window (ksw.c ksw_align2 with KSW_XSUBO | KSW_XSTART | minsc):
 * score, end (te, qe inclusive), second-best score away from the best (score2, te2; -1 = none) and start (tb, qb; -1 when
 * score < minsc).  Returns executed cells. */
typedef struct { int32_t score, te, qe, score2, te2, tb, qb, pad; } qmo_sw_t;
int64_t qmo_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const qmo_opt_t *o,
                       int minsc, qmo_sw_t *r);

/* ---- reference + k-mer index (replaces `bwa index`, rules/index.smk:13) ---- */
#define QMO_MAX_SEEDS 64
#define QMO_MAX_REGS  16
#define QMO_MAX_CIGAR 21
#define QMO_OCC_CAP   32
#define QMO_MAX_MATESW 50            /* bwa's max_matesw (never reached: at most QMO_MAX_REGS anchors per end) */
#define QMO_RESCUE_MAX_WINDOW 4096   /* mate-rescue windows longer than this are not searched                   */

typedef struct qmo_ref qmo_ref_t;
/* codes: concatenated forward strands (0..3), contig i has lens[i] bases */
qmo_ref_t *qmo_ref_create(const uint8_t *codes, int n_contigs, const int64_t *lens, int k);
void qmo_ref_destroy(qmo_ref_t *r);
int64_t qmo_ref_lpac(const qmo_ref_t *r);
void qmo_ref_set_fm(qmo_ref_t *r, const void *fm /* qmo_fm_t, not owned */, int max_mem_intv);

typedef struct { int64_t rbeg; int32_t qbeg, len; } qmo_seed_t;              /* 16 B */
typedef struct {
    int64_t rb, re;               /* [rb,re) in bwa's doubled coordinates (>= l_pac: reverse strand) */
    int32_t qb, qe;
    int32_t rid, score, truesc, sub, csub, sub_n, w, seedcov, secondary, seedlen0;
} qmo_reg_t;                                                                   /* 64 B */
typedef struct {
    int32_t rid, pos;             /* 0-based leftmost position on contig rid                         */
    uint16_t flag; uint8_t mapq; uint8_t n_cigar;
    int32_t score, sub, nm;
    int32_t mate_rid, mate_pos, tlen;
    int32_t qb, qe;               /* aligned interval of the read as sequenced                        */
    uint32_t cigar[QMO_MAX_CIGAR];/* BAM encoding len<<4|op, ops M=0 I=1 D=2 S=4                       */
} qmo_aln_t;                                                                   /* 128 B */
typedef struct { int32_t low, high, failed, pad; double avg, std; } qmo_pestat_t;   /* x4: FF FR RF RR */

/* log of every ksw_extend2 call made while aligning (stage-local parity: "same seeds => same tuples") */
typedef struct {
    uint32_t q_off, t_off; int32_t qlen, tlen, h0, w, end_bonus; uint32_t flags;   /* = qm_ext_task */
} qmo_ext_task_t;
typedef struct {
    uint8_t *seq; int64_t seq_len, seq_cap;
    qmo_ext_task_t *tasks; qmo_ext_t *results; int32_t *w_used; int64_t *cells; int64_t n, cap;
} qmo_ext_log_t;

/* MEM seeds of one read (codes 0..4), sorted by (qbeg, rbeg); returns count */
int qmo_collect_seeds(const qmo_ref_t *R, const qmo_opt_t *o, const uint8_t *read, int len, qmo_seed_t *out);

/* single-end stage for reads [0,n): seeds -> chains -> extension -> sorted, de-duplicated regions
 * (bwamem.c mem_align1_core).  reads: n x stride codes.  Any of the outputs may be NULL. */
void qmo_align_se(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n, const uint8_t *reads, int stride,
                  const int32_t *lens, qmo_seed_t *seeds, int32_t *n_seeds, qmo_reg_t *regs, int32_t *n_regs,
                  qmo_ext_log_t *log, int64_t *cells_total);

/* insert-size model from the first n_pairs pairs' regions (bwamem_pair.c mem_pestat) */
void qmo_pestat(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, const qmo_reg_t *regs,
                const int32_t *n_regs, qmo_pestat_t pes[4]);

/* mate rescue for every pair (bwamem_pair.c mem_matesw as driven by mem_sam_pe): regs / n_regs updated in place;
 * n_sw / cells (may be NULL) = local alignments run and their cells */
void qmo_mate_rescue(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, const uint8_t *reads, int stride,
                     const int32_t *lens, qmo_reg_t *regs, int32_t *n_regs, const qmo_pestat_t pes[4],
                     int64_t *n_sw_total, int64_t *cells_total);

/* paired-end stage (bwamem_pair.c mem_sam_pe; mate rescue unless flags & QMO_F_NO_RESCUE) + CIGAR generation
 * (bwamem.c mem_reg2aln / bwa.c bwa_gen_cigar2).  pair_id0 = global index of pair 0 (tie-break hash).
 * regs/n_regs are modified in place (primary marking re-sorts them). */
void qmo_pair_and_finish(const qmo_ref_t *R, const qmo_opt_t *o, int64_t n_pairs, int64_t pair_id0,
                         const uint8_t *reads, int stride, const int32_t *lens,
                         qmo_reg_t *regs, int32_t *n_regs, const qmo_pestat_t pes[4], qmo_aln_t *alns);

/* ---- pileup counting (bcftools mpileup -B semantics, rules/vcfcall.smk:115) ----
 * counts: int32 [l_pac][16], row = forward reference position (contigs concatenated), channels:
 *  0-3 A,C,G,T forward (BQ >= min_bq)   4 N forward   5 deleted forward
 *  6-9 A,C,G,T reverse                 10 N reverse  11 deleted reverse
 *  12 insertion follows this base  13 deletion follows this base
 *  14 raw depth (every non-deleted base of every admitted read, before the BQ filter = bcftools ori_depth)
 *  15 admitted reads whose leftmost aligned base is here */
#define QMO_NCH 16
typedef struct {
    int32_t min_mapq;        /* -q (0)   */
    int32_t min_bq;          /* -Q (13)  */
    int32_t count_orphans;   /* -A       */
    int32_t ignore_overlaps; /* -x       */
} qmo_pileup_opt_t;
void qmo_pileup_opt_default(qmo_pileup_opt_t *p);
void qmo_pileup(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_pairs, const qmo_aln_t *alns,
                const uint8_t *reads, const uint8_t *quals, int stride, const int32_t *lens, int32_t *counts);

/* samtools-mpileup text of the batch (rules/vcfcall.smk:39) with the admission / overlap rules above;
 * names[i] = name of contig i.  *out is malloc'ed (release with qmo_free); returns its length (< 2 GB). */
int64_t qmo_mpileup_text(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_pairs, const qmo_aln_t *alns,
                         const uint8_t *reads, const uint8_t *quals, int stride, const int32_t *lens,
                         const char *const *names, char **out);
void qmo_free(void *p);

/* ---- base alignment quality (htslib 1.9 realn.c sam_prob_realn / probaln.c kpa_glocal, what both mpileups run unless -B is
 * given: rules/vcfcall.smk:39,115 give no -B).  PARITY UNPINNED (no htslib here): qmo_baq.c. ----
 * quals_out = quals (reads as sequenced) with every admitted read's base qualities capped by its BAQ; flag 3 = extended BAQ
 * (what the mpileups pass), 1 = plain. */
void qmo_baq(const qmo_ref_t *R, const qmo_pileup_opt_t *po, int64_t n_reads, const qmo_aln_t *alns, const uint8_t *reads,
             const uint8_t *quals, int stride, const int32_t *lens, int flag, uint8_t *quals_out);
/* the HMM alone: ref / query base codes (0..3, above = ambiguous), iqual phred values; state[i] = reference offset << 2 | (0 M, 1 I),
 * q[i] = phred of the posterior of that state; returns the phred-scaled likelihood */
int qmo_kpa_glocal(const uint8_t *ref, int l_ref, const uint8_t *query, int l_query, const uint8_t *iqual,
                   double d, double e, int bw, int *state, uint8_t *q);

#ifdef __cplusplus
}
#endif
#endif
