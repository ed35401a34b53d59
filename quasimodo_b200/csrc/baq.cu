// baq.cu -- base alignment quality: the per-base quality cap `bcftools mpileup` and `samtools mpileup` apply to every read
// unless -B is given (reference call sites rules/vcfcall.smk:39,115 pass no -B).  Upstream: htslib 1.9
// realn.c sam_prob_realn(flag 3 = apply + extended) over probaln.c kpa_glocal -- a banded profile HMM (states M / I / D per
// query base x reference base, forward / backward with per-row rescaling, posterior decoding) in DOUBLE precision whose
// posteriors are rounded to phred integers.  Off by default (the north-star parity configuration is -B); qm_baq_apply /
// qm_sample_set_baq turn it on.  Oracle: oracle/qmo_baq.c (restated from the published algorithm; parity unpinned).
//
// Bit-exactness with the CPU restatement needs the same IEEE operations in the same order: the library is built with
// -fmad=false (no contraction into FMAs), double division and the float error table are IEEE on both sides, and the HMM below
// keeps upstream's statement order.  log() is the one library call: CUDA's is within 1 ulp of glibc's, and a difference only
// shows when -4.343 ln(1 - p) + .499 sits within ~1e-15 of an integer.
//
// Layout on the GPU: one THREAD per read (a read's HMM is a serial recurrence over 150 x 15 x 3 cells; reads are the
// parallelism).  Reads are sorted into two classes by band width:
//   bw == 7 (no net indel above 7 bases: all but a few reads)  rows of 51 doubles, scratch interleaved [cell][thread] so a
//                                                              warp's accesses coalesce; a chunk of T reads per launch
//   wider bands                                                a contiguous slab per read, as many reads per launch as fit
// The forward matrix is kept (61 KB per 150-base read); the backward pass keeps two rows and decodes each row as it is made.
#include <math.h>
#include <vector>
#include "pipeline.cuh"
#include "baq_core.cuh"

namespace {

__constant__ float c_qual2prob[256];

constexpr int kFastBw = 7, kFastRow = (2 * kFastBw + 1) * 3 + 6;
constexpr size_t kBaqScratchBudget = (size_t)6 << 30;

__global__ void __launch_bounds__(256)
baq_prep_kernel(IndexView V, qm_pileup_opt po, const qm_aln *__restrict__ alns, const int32_t *__restrict__ lens, int64_t n,
                BaqRead *__restrict__ fast, BaqRead *__restrict__ slow, unsigned long long *__restrict__ slow_cells, int *__restrict__ ctr)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n) return;
    const qm_aln a = alns[r];
    if (!baq_admitted(po, a)) return;
    const int L = lens[r];
    int xb, xe, bw;
    if (!baq_window(a, L, V.len[a.rid], xb, xe, bw)) return;
    const int hb = baq_band(xe - xb, L, bw);
    BaqRead br = {(int32_t)r, xb, xe, bw};
    if (hb == kFastBw) fast[atomicAdd(&ctr[0], 1)] = br;
    else {
        const int k = atomicAdd(&ctr[1], 1);
        slow[k] = br;
        slow_cells[k] = (unsigned long long)(L + 3) * ((2 * hb + 1) * 3 + 6) + (L + 2);      // doubles of the read's slab
    }
}

// scratch accessors: f(i, c) forward rows, b(j, c) two backward rows, s(i) scaling factors
struct FastAcc {
    double *base; size_t T, t; int L;
    __device__ __forceinline__ double &f(int i, int c) const { return base[((size_t)i * kFastRow + c) * T + t]; }
    __device__ __forceinline__ double &b(int j, int c) const { return base[((size_t)(L + 1 + j) * kFastRow + c) * T + t]; }
    __device__ __forceinline__ double &s(int i) const { return base[((size_t)(L + 3) * kFastRow + i) * T + t]; }
};
struct SlowAcc {
    double *base; int row, L;
    __device__ __forceinline__ double &f(int i, int c) const { return base[(size_t)i * row + c]; }
    __device__ __forceinline__ double &b(int j, int c) const { return base[(size_t)(L + 1 + j) * row + c]; }
    __device__ __forceinline__ double &s(int i) const { return base[(size_t)(L + 3) * row + i]; }
};

__global__ void __launch_bounds__(128)
baq_fast_kernel(IndexView V, const BaqRead *__restrict__ list, int n, const qm_aln *__restrict__ alns, const uint8_t *__restrict__ codes,
                const uint8_t *__restrict__ quals, int stride, const int32_t *__restrict__ lens, int flag, uint8_t *__restrict__ quals_out,
                double *__restrict__ work, int32_t *__restrict__ st, uint8_t *__restrict__ qv, int T, int Lmax)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const BaqRead br = list[t];
    const int L = lens[br.read];
    FastAcc A = {work, (size_t)T, (size_t)t, Lmax};
    const qm_aln &a = alns[br.read];
    baq_one(V.refb + V.off[a.rid] + br.xb, c_qual2prob, A, br, a, codes, quals, stride, L, flag, quals_out, st + t, qv + t, (size_t)T);
}

__global__ void __launch_bounds__(128)
baq_slow_kernel(IndexView V, const BaqRead *__restrict__ list, const unsigned long long *__restrict__ offs, int n, const qm_aln *__restrict__ alns,
                const uint8_t *__restrict__ codes, const uint8_t *__restrict__ quals, int stride, const int32_t *__restrict__ lens, int flag,
                uint8_t *__restrict__ quals_out, double *__restrict__ work, int32_t *__restrict__ st, uint8_t *__restrict__ qv, int Lmax)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const BaqRead br = list[t];
    const int L = lens[br.read];
    const int hb = baq_band(br.xe - br.xb, L, br.bw);
    SlowAcc A = {work + offs[t], (2 * hb + 1) * 3 + 6, L};
    const qm_aln &a = alns[br.read];
    baq_one(V.refb + V.off[a.rid] + br.xb, c_qual2prob, A, br, a, codes, quals, stride, L, flag, quals_out, st + (size_t)t * Lmax, qv + (size_t)t * Lmax, (size_t)1);
}

}  // namespace

extern "C" {

int qm_baq_apply(qm_ctx *ctx, const qm_index *idx, const qm_pileup_opt *po, const qm_aln *d_alns, const uint8_t *d_codes,
                 const uint8_t *d_quals, int32_t stride, const int32_t *d_lens, int64_t n_reads, int32_t flag, uint8_t *d_quals_out,
                 void *stream)
{
    if (!ctx || !idx || !po || n_reads < 0 || stride <= 0 || (n_reads > 0 && (!d_alns || !d_codes || !d_quals || !d_lens || !d_quals_out)))
        return QM_EINVAL;
    if (!(flag & 1)) return qm_fail(ctx, QM_EINVAL, "qm_baq_apply: flag must have bit 0 set (1 = BAQ, 3 = extended BAQ)");
    if (n_reads > 0x7fffffff) return qm_fail(ctx, QM_ELIMIT, "qm_baq_apply: at most 2^31-1 reads per call");
    if (n_reads == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    static bool table_set[64] = {};
    if (!table_set[ctx->device & 63]) {
        float tab[256];
        for (int i = 0; i < 256; ++i) tab[i] = (float)pow(10, -i / 10.);          // g_qual2prob, stored as float like upstream's qual[]
        QM_CUDA(ctx, cudaMemcpyToSymbol(c_qual2prob, tab, sizeof(tab)));
        table_set[ctx->device & 63] = true;
    }
    if (d_quals_out != d_quals) QM_CUDA(ctx, cudaMemcpyAsync(d_quals_out, d_quals, (size_t)n_reads * stride, cudaMemcpyDeviceToDevice, st));
    // lists of the two classes
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_fast = 256, o_slow = o_fast + al((size_t)n_reads * sizeof(BaqRead)), o_cells = o_slow + al((size_t)n_reads * sizeof(BaqRead));
    const size_t lists_bytes = o_cells + al((size_t)n_reads * 8);
    void *lp = nullptr;
    int rc = qm_scratch_reserve(ctx, 27, lists_bytes, &lp);
    if (rc) return rc;
    int *d_ctr = (int *)lp;
    BaqRead *d_fast = (BaqRead *)((char *)lp + o_fast), *d_slow = (BaqRead *)((char *)lp + o_slow);
    unsigned long long *d_cells = (unsigned long long *)((char *)lp + o_cells);
    QM_CUDA(ctx, cudaMemsetAsync(d_ctr, 0, 256, st));
    baq_prep_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(idx->v, *po, d_alns, d_lens, n_reads, d_fast, d_slow, d_cells, d_ctr);
    int h_ctr[2] = {0, 0};
    QM_CUDA(ctx, cudaMemcpyAsync(h_ctr, d_ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    const int Lmax = stride;
    // fast class: chunks of T reads, T sized by the scratch budget
    const size_t per_read = ((size_t)(Lmax + 3) * kFastRow + (Lmax + 2)) * 8 + (size_t)Lmax * 5;
    if (h_ctr[0] > 0) {
        size_t T = kBaqScratchBudget / per_read;
        T = T > (size_t)h_ctr[0] ? (size_t)h_ctr[0] : T;
        T = (T + 127) & ~(size_t)127;
        const size_t o_st = al(T * ((size_t)(Lmax + 3) * kFastRow + (Lmax + 2)) * 8), o_qv = o_st + al(T * Lmax * 4);
        void *wp = nullptr;
        rc = qm_scratch_reserve(ctx, 28, o_qv + al(T * Lmax), &wp);
        if (rc) return rc;
        for (int64_t c0 = 0; c0 < h_ctr[0]; c0 += (int64_t)T) {
            const int nn = (int)(h_ctr[0] - c0 < (int64_t)T ? h_ctr[0] - c0 : (int64_t)T);
            baq_fast_kernel<<<(nn + 127) / 128, 128, 0, st>>>(idx->v, d_fast + c0, nn, d_alns, d_codes, d_quals, stride, d_lens, flag, d_quals_out,
                                                              (double *)wp, (int32_t *)((char *)wp + o_st), (uint8_t *)((char *)wp + o_qv), (int)T, Lmax);
        }
    }
    // wide bands: slabs sized per read (prefix sums on the host: a few reads)
    if (h_ctr[1] > 0) {
        std::vector<unsigned long long> cells(h_ctr[1]), offs(h_ctr[1]);
        QM_CUDA(ctx, cudaMemcpyAsync(cells.data(), d_cells, (size_t)h_ctr[1] * 8, cudaMemcpyDeviceToHost, st));
        QM_CUDA(ctx, cudaStreamSynchronize(st));
        int c0 = 0;
        while (c0 < h_ctr[1]) {
            unsigned long long acc = 0;
            int c1 = c0;
            while (c1 < h_ctr[1] && (c1 == c0 || (acc + cells[c1]) * 8 <= kBaqScratchBudget)) { offs[c1] = acc; acc += cells[c1]; ++c1; }
            const int nn = c1 - c0;
            const size_t o_offs = al((size_t)acc * 8), o_st = o_offs + al((size_t)nn * 8), o_qv = o_st + al((size_t)nn * Lmax * 4);
            void *wp = nullptr;
            rc = qm_scratch_reserve(ctx, 28, o_qv + al((size_t)nn * Lmax), &wp);
            if (rc) return rc;
            QM_CUDA(ctx, cudaMemcpyAsync((char *)wp + o_offs, offs.data() + c0, (size_t)nn * 8, cudaMemcpyHostToDevice, st));
            baq_slow_kernel<<<(nn + 127) / 128, 128, 0, st>>>(idx->v, d_slow + c0, (const unsigned long long *)((char *)wp + o_offs), nn, d_alns, d_codes,
                                                              d_quals, stride, d_lens, flag, d_quals_out, (double *)wp,
                                                              (int32_t *)((char *)wp + o_st), (uint8_t *)((char *)wp + o_qv), Lmax);
            QM_CUDA(ctx, cudaStreamSynchronize(st));          // offs[] is reused by the next group
            c0 = c1;
        }
    }
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

int qm_baq_apply_host(qm_ctx *ctx, const qm_index *idx, const qm_pileup_opt *po, const qm_aln *h_alns, const uint8_t *h_codes,
                      const uint8_t *h_quals, int32_t stride, const int32_t *h_lens, int64_t n_reads, int32_t flag, uint8_t *h_quals_out)
{
    if (!ctx || !idx || !po || n_reads < 0 || stride <= 0 || (n_reads > 0 && (!h_alns || !h_codes || !h_quals || !h_lens || !h_quals_out)))
        return QM_EINVAL;
    if (n_reads == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n_reads * stride;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_q = al(sb), o_o = o_q + al(sb), o_l = o_o + al(sb), o_a = o_l + al((size_t)n_reads * 4);
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 29, o_a + al((size_t)n_reads * sizeof(qm_aln)), &p);
    if (rc) return rc;
    char *b = (char *)p;
    cudaStream_t st = ctx->own_stream;
    QM_CUDA(ctx, cudaMemcpyAsync(b, h_codes, sb, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_q, h_quals, sb, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_l, h_lens, (size_t)n_reads * 4, cudaMemcpyHostToDevice, st));
    QM_CUDA(ctx, cudaMemcpyAsync(b + o_a, h_alns, (size_t)n_reads * sizeof(qm_aln), cudaMemcpyHostToDevice, st));
    rc = qm_baq_apply(ctx, idx, po, (const qm_aln *)(b + o_a), (const uint8_t *)b, (const uint8_t *)(b + o_q), stride, (const int32_t *)(b + o_l),
                      n_reads, flag, (uint8_t *)(b + o_o), st);
    if (rc) return rc;
    QM_CUDA(ctx, cudaMemcpyAsync(h_quals_out, b + o_o, sb, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    return QM_OK;
}

}  // extern "C"
