/*
 * quasimodo_b200.h -- C-ABI of the B200-native QuasiModo read-level hot path.
 *
 * The reference (hzi-bifo/Quasimodo) has NO in-process plugin/FFI interface: its hot path is a chain of
 * Snakemake `shell:` lines calling external binaries.  Each entry point below
 * therefore cites the rule line / upstream function whose work it replaces.  Conventions:
 *   - extern "C", plain pointers and sizes, no C++/torch types;
 *   - every function returns 0 on success or a negative QM_E* code; qm_last_error(ctx) explains;
 *   - pointers named d_* are DEVICE pointers owned by the caller, h_* are HOST pointers owned by the
 *     caller; the library never frees caller memory and never returns memory the caller must free;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t. the host unless their
 *     name ends in _host or _sync;
 *   - a qm_ctx is bound to one CUDA device and is not thread-safe; distinct ctxs are independent;
 *   - there is NO CPU fallback: without a CUDA device qm_ctx_create fails with QM_ENODEV.
 */
#ifndef QUASIMODO_B200_H
#define QUASIMODO_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QM_OK        0
#define QM_EINVAL   -1   /* bad argument                                   */
#define QM_ENODEV   -2   /* no usable CUDA device / extension not usable   */
#define QM_ECUDA    -3   /* CUDA runtime error (see qm_last_error)         */
#define QM_ENOMEM   -4
#define QM_ELIMIT   -5   /* input exceeds a documented hard limit          */
#define QM_EIO      -6

typedef struct qm_ctx qm_ctx;

/* ---- scoring and bwa-mem options in effect for `bwa mem -k 31` (rules/bwa.smk:15) ---- */
typedef struct {
    int32_t a, b;                 /* match score, mismatch penalty (N vs anything = -1)  */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t w;                    /* band width (-w)                                     */
    int32_t zdrop;                /* -d                                                  */
    int32_t pen_clip5, pen_clip3; /* -L                                                  */
    int32_t min_seed_len;         /* -k                                                  */
    int32_t max_occ;              /* -c                                                  */
    int32_t T;                    /* -T                                                  */
    int32_t pen_unpaired;         /* -U                                                  */
    int32_t max_ins;
    int32_t max_chain_gap;
    int32_t mapq_coef_len;
    float   mask_level, drop_ratio, mask_level_redun;
    int32_t min_chain_weight;
    int32_t reserved[3];
} qm_opt;

void qm_opt_default(qm_opt *opt);

/* ---- context ---- */
int         qm_ctx_create(int device, qm_ctx **out);
void        qm_ctx_destroy(qm_ctx *ctx);
const char *qm_last_error(const qm_ctx *ctx);          /* owned by ctx */
const char *qm_version(void);
int         qm_device_sm_count(const qm_ctx *ctx);

/* ---- alignment: batched banded affine-gap extension ----
 * Replaces bwa ksw.c:ksw_extend2 as called from bwamem.c:mem_chain2aln (reference call site
 * rules/bwa.smk:15 `bwa mem -k 31`); semantics in SURVEY.md Appendix A.3.  This is the parity entry
 * point: one task = one ksw_extend2 call.  Sequences are base codes 0..4 (A C G T N), one byte each,
 * in one device arena `d_seq`; a task addresses its query/target by byte offset.
 * If (flags & QM_EXT_BAND_RETRY) the task is run as mem_chain2aln runs it: with w, then with 2w when
 * the first try's score moved and max_off >= 3/4 w (MAX_BAND_TRY = 2); w_used reports the last band.
 */
#define QM_EXT_BAND_RETRY 1u
#define QM_EXT_PREV_H0    2u       /* retry loop starts with prev = h0 (right extension: prev = sc0),
                                      otherwise prev = -1 (left extension)                            */
#define QM_EXT_MAX_QLEN   511      /* hard limit of the kernel's column striping */

typedef struct {
    uint32_t q_off, t_off;        /* byte offsets into d_seq                         */
    int32_t  qlen, tlen;
    int32_t  h0;                  /* score of the seed (+ left extension)            */
    int32_t  w;                   /* band width for this call                        */
    int32_t  end_bonus;           /* pen_clip5 / pen_clip3                           */
    uint32_t flags;
} qm_ext_task;                    /* 32 bytes */

typedef struct {
    int32_t score, qle, tle, gtle, gscore, max_off;   /* ksw_extend2's return value and out-params */
    int32_t w_used;               /* band of the last try                            */
    int32_t cells;                /* inner-loop cells the reference loop executes (sum of end-beg
                                     over rows, over all tries) -- the GCUPS work unit */
} qm_ext_result;                  /* 32 bytes */

int qm_extend_batch(qm_ctx *ctx, const qm_opt *opt, const uint8_t *d_seq, const qm_ext_task *d_tasks,
                    int64_t n_tasks, qm_ext_result *d_out, void *stream);

/* same, host buffers: copies seq + tasks in, results out, synchronises.  seq_bytes = arena size. */
int qm_extend_batch_host(qm_ctx *ctx, const qm_opt *opt, const uint8_t *h_seq, size_t seq_bytes,
                         const qm_ext_task *h_tasks, int64_t n_tasks, qm_ext_result *h_out);

/* ---- synthetic inputs: deterministic, index-addressable read-pair simulator ----
 * The reference ships no reads (data/PRJEB32127.txt lists ENA URLs; no network), so benchmark and
 * parity inputs are simulated from the bundled genomes.  Pair i is a pure function of (seed, i):
 * any shard can be generated independently, on the host or on the device, with identical bytes.
 * genome: base codes 0..3 of all source genomes, source s at [src_off[s], src_off[s]+src_len[s]);
 * src_cum[s] = floor(2^32 * cumulative weight share) (last = 2^32-1), weight = copies x length.
 * Output layout ("read batch", used by every later stage): reads 2i / 2i+1 are the mates of pair i,
 * codes[r*stride + j] in 0..4 (4 = N), quals[r*stride + j] = phred, stride >= read_len. */
typedef struct {
    uint64_t seed;
    int32_t  read_len;
    int32_t  ins_mean, ins_sd, ins_max;   /* insert size model N(mean, sd) clipped to [read_len, max] */
    int32_t  n_sources;
    int32_t  indel_ppm;                   /* indel events per 1e6 bases (cfg 5: 200)                  */
    int32_t  n_ppm;                       /* bases forced to N per 1e6 (1000)                         */
    int32_t  lowq_ppm;                    /* bases with Q2..Q12 per 1e6 (20000)                       */
    int32_t  reserved[3];
} qm_sim_params;

int qm_simulate_pairs_host(const qm_sim_params *p, const uint8_t *h_genome, const int64_t *src_off,
                           const int64_t *src_len, const uint32_t *src_cum, int64_t pair0, int64_t n_pairs,
                           int32_t stride, uint8_t *h_codes, uint8_t *h_quals, int32_t *h_src /* may be NULL */,
                           int64_t *h_pos /* may be NULL */);
int qm_simulate_pairs(qm_ctx *ctx, const qm_sim_params *p, const uint8_t *d_genome, const int64_t *h_src_off,
                      const int64_t *h_src_len, const uint32_t *h_src_cum, int64_t pair0, int64_t n_pairs,
                      int32_t stride, uint8_t *d_codes, uint8_t *d_quals, void *stream);

/* ---- measurement helper: whole-GPU issue rate of the DPX instruction the extension kernel leans on
 * (register-resident dependent chains of __viaddmax_s16x2_relu / __viaddmax_s32_relu on every SM).
 * Returns giga warp-lane instructions per second in *out_gops (packed: per 32-bit lane op).
 * kind: 0 = __viaddmax_s32_relu, 1 = __viaddmax_s16x2_relu, 2 = __vimax3_s32.  SURVEY.md 8d / BASELINE.md 2. */
int qm_dpx_peak_sync(qm_ctx *ctx, int kind, int iters, double *out_gops, double *out_ms);

#ifdef __cplusplus
}
#endif
#endif /* QUASIMODO_B200_H */
