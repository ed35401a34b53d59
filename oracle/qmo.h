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
 *     follows the published algorithms and is cross-checked against an
 *     independent Python restatement (oracle/ksw_py.py).
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
    int32_t reserved[3];
} qmo_opt_t;

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

#ifdef __cplusplus
}
#endif
#endif
