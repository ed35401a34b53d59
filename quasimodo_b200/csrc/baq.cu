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

namespace {

__constant__ float c_qual2prob[256];

constexpr double kEI = .25, kEM = .33333333333;
constexpr int kFastBw = 7, kFastRow = (2 * kFastBw + 1) * 3 + 6;
constexpr size_t kBaqScratchBudget = (size_t)6 << 30;

struct BaqRead { int32_t read, xb, xe, bw; };

__device__ __forceinline__ bool baq_admitted(const qm_pileup_opt &po, const qm_aln &a)
{
    if (a.flag & (0x4 | 0x100 | 0x200 | 0x400)) return false;
    if (a.n_cigar == 0 || a.n_cigar == 255) return false;
    if ((int)a.mapq < po.min_mapq) return false;
    if ((a.flag & 0x1) && !(a.flag & 0x2) && !po.count_orphans) return false;
    return true;
}

// the reference window and band of one read (sam_prob_realn's first half); false = the read is left alone
__device__ bool baq_window(const qm_aln &a, int l_qseq, int64_t ref_len, int &xb_o, int &xe_o, int &bw_o)
{
    int x = a.pos, y = 0, yb = -1, ye = -1, xb = -1, xe = -1;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, l = (int)(a.cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (yb < 0) yb = y;
            if (xb < 0) xb = x;
            ye = y + l; xe = x + l;
            x += l; y += l;
        } else if (op == 4 || op == 1) y += l;
        else if (op == 2) x += l;
        else if (op == 3) return false;
    }
    if (xb < 0 || l_qseq <= 0) return false;
    int bw = 7;
    const int dd = abs((xe - xb) - (ye - yb));
    if (dd > bw) bw = dd + 3;
    xb -= yb + bw / 2; if (xb < 0) xb = 0;
    xe += l_qseq - ye + bw / 2;
    if (xe - xb - l_qseq > bw) {
        xb += (xe - xb - l_qseq - bw) / 2;
        xe -= (xe - xb - l_qseq - bw) / 2;               // sees the line above's xb, as upstream
    }
    if (xe > ref_len) xe = (int)ref_len;
    if (xe - xb <= 0) return false;
    xb_o = xb; xe_o = xe; bw_o = bw;
    return true;
}

// the band the HMM really uses (kpa_glocal's first lines)
__device__ __forceinline__ int baq_band(int l_ref, int l_query, int cbw)
{
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > cbw) bw = cbw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    return bw;
}

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

#define SET_U(u, b, i, k) { int x_ = (i) - (b); x_ = x_ > 0 ? x_ : 0; (u) = ((k) - x_ + 1) * 3; }

// probaln.c kpa_glocal on the scratch of accessor A.  ref(k), qry(i), err(i): 1-based reference code, query code, error
// probability (float) of the read in BAM orientation.  out_state / out_q (i = 0 .. l_query-1) receive the decoding.
template <class Acc, class Ref, class Qry, class Err, class Out>
__device__ void baq_glocal(const Acc &A, int l_ref, int l_query, int cbw, double cd, double ce, Ref ref, Qry qry, Err err, Out out)
{
    const int bw = baq_band(l_ref, l_query, cbw);
    const int bw2 = bw * 2 + 1, row = bw2 * 3 + 6;
    double m[9];
    const double sM = 1. / (2 * l_query + 2), sI = sM;
    m[0] = (1 - cd - cd) * (1 - sM); m[1] = m[2] = cd * (1 - sM);
    m[3] = (1 - ce) * (1 - sI); m[4] = ce * (1 - sI); m[5] = 0.;
    m[6] = 1 - ce; m[7] = 0.; m[8] = ce;
    const double bM = (1 - cd) / l_ref, bI = cd / l_ref;
    // every cell the recurrences may look at starts as zero (upstream: calloc)
    for (int i = 0; i <= l_query; ++i) for (int c = 0; c < row; ++c) A.f(i, c) = 0.;
    for (int j = 0; j < 2; ++j) for (int c = 0; c < row; ++c) A.b(j, c) = 0.;
    int k;
    /*** forward ***/
    SET_U(k, bw, 0, 0);
    A.f(0, k) = 1.; A.s(0) = 1.;
    {
        double sum = 0.;
        const int beg = 1, end = l_ref < bw + 1 ? l_ref : bw + 1;
        const int q1 = qry(1);
        const double ql = (double)err(1);
        for (k = beg; k <= end; ++k) {
            int u;
            const int rk = ref(k);
            const double e = (rk > 3 || q1 > 3) ? 1. : rk == q1 ? 1. - ql : ql * kEM;
            SET_U(u, bw, 1, k);
            const double f0 = e * bM, f1 = kEI * bI;
            A.f(1, u) = f0; A.f(1, u + 1) = f1;
            sum += f0 + f1;
        }
        A.s(1) = sum;
        int _beg, _end;
        SET_U(_beg, bw, 1, beg); SET_U(_end, bw, 1, end); _end += 2;
        for (k = _beg; k <= _end; ++k) A.f(1, k) = A.f(1, k) / sum;
    }
    for (int i = 2; i <= l_query; ++i) {
        double sum = 0.;
        const double qli = (double)err(i);
        int beg = 1, end = l_ref, x, _beg, _end;
        const int qyi = qry(i);
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg; k <= end; ++k) {
            int u, v11, v01, v10;
            const int rk = ref(k);
            const double e = (rk > 3 || qyi > 3) ? 1. : rk == qyi ? 1. - qli : qli * kEM;
            SET_U(u, bw, i, k); SET_U(v11, bw, i - 1, k - 1); SET_U(v10, bw, i - 1, k); SET_U(v01, bw, i, k - 1);
            const double f0 = e * (m[0] * A.f(i - 1, v11) + m[3] * A.f(i - 1, v11 + 1) + m[6] * A.f(i - 1, v11 + 2));
            const double f1 = kEI * (m[1] * A.f(i - 1, v10) + m[4] * A.f(i - 1, v10 + 1));
            const double f2 = m[2] * A.f(i, v01) + m[8] * A.f(i, v01 + 2);
            A.f(i, u) = f0; A.f(i, u + 1) = f1; A.f(i, u + 2) = f2;
            sum += f0 + f1 + f2;
        }
        A.s(i) = sum;
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        sum = 1. / sum;
        for (k = _beg; k <= _end; ++k) A.f(i, k) = A.f(i, k) * sum;
    }
    {
        double sum = 0.;
        for (k = 1; k <= l_ref; ++k) {
            int u;
            SET_U(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += A.f(l_query, u) * sM + A.f(l_query, u + 1) * sI;
        }
        A.s(l_query + 1) = sum;
    }
    /*** backward + posterior decoding, row by row: row i lives in b(i & 1) ***/
    auto decode = [&](int i) {
        double sum = 0., mx = 0.;
        int beg = 1, end = l_ref, x, max_k = -1;
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (int kk = beg; kk <= end; ++kk) {
            int u;
            SET_U(u, bw, i, kk);
            double z = A.f(i, u) * A.b(i & 1, u); if (z > mx) { mx = z; max_k = (kk - 1) << 2 | 0; } sum += z;
            z = A.f(i, u + 1) * A.b(i & 1, u + 1); if (z > mx) { mx = z; max_k = (kk - 1) << 2 | 1; } sum += z;
        }
        mx /= sum;
        int qq = (int)(-4.343 * log(1. - mx) + .499);
        out(i - 1, max_k, qq > 100 ? 99 : qq);
    };
    {
        const int j = l_query & 1;
        const double sl = A.s(l_query), sl1 = A.s(l_query + 1);
        for (k = 1; k <= l_ref; ++k) {
            int u;
            SET_U(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            A.b(j, u) = sM / sl / sl1; A.b(j, u + 1) = sI / sl / sl1;
        }
        decode(l_query);
    }
    for (int i = l_query - 1; i >= 1; --i) {
        const int j = i & 1, j1 = j ^ 1;
        for (int c = 0; c < row; ++c) A.b(j, c) = 0.;                   // this slot held row i + 2
        int beg = 1, end = l_ref, x, _beg, _end;
        double y = (i > 1);
        const double qli1 = (double)err(i + 1);
        const int qyi1 = qry(i + 1);
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = end; k >= beg; --k) {
            int u, v11, v01, v10;
            SET_U(u, bw, i, k); SET_U(v11, bw, i + 1, k + 1); SET_U(v10, bw, i + 1, k); SET_U(v01, bw, i, k + 1);
            double e;
            if (k >= l_ref) e = 0;
            else { const int rk = ref(k + 1); e = (rk > 3 || qyi1 > 3) ? 1. : rk == qyi1 ? 1. - qli1 : qli1 * kEM; }
            e = e * A.b(j1, v11);
            const double b10 = A.b(j1, v10 + 1), b01 = A.b(j, v01 + 2);
            A.b(j, u) = e * m[0] + kEI * m[1] * b10 + m[2] * b01;
            A.b(j, u + 1) = e * m[3] + kEI * m[4] * b10;
            A.b(j, u + 2) = (e * m[6] + m[8] * b01) * y;
        }
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        y = 1. / A.s(i);
        for (k = _beg; k <= _end; ++k) A.b(j, k) = A.b(j, k) * y;
        decode(i);
    }
}

// One read: window, HMM, then sam_prob_realn's second half (agreement with the CIGAR, extended BAQ inside each M block, the
// cap).  st / qv: per-read scratch for the decoding (state int32, phred byte), strided by `ss`.
template <class Acc>
__device__ void baq_one(const IndexView &V, const Acc &A, const BaqRead br, const qm_aln &a, const uint8_t *__restrict__ codes,
                        const uint8_t *__restrict__ quals, int stride, int L, int flag, uint8_t *__restrict__ quals_out,
                        int32_t *st, uint8_t *qv, size_t ss)
{
    const bool rev = (a.flag & 0x10) != 0;
    const uint8_t *rd = codes + (size_t)br.read * stride, *ql = quals + (size_t)br.read * stride;
    uint8_t *qo = quals_out + (size_t)br.read * stride;
    const uint8_t *refw = V.refb + V.off[a.rid] + br.xb;
    auto ref = [&](int k) { const int c = refw[k - 1]; return c > 3 ? 4 : c; };
    auto qry = [&](int i) { const int c = rev ? rd[L - i] : rd[i - 1]; return c > 3 ? 4 : (rev ? 3 - c : c); };
    auto err = [&](int i) { return c_qual2prob[rev ? ql[L - i] : ql[i - 1]]; };
    auto out = [&](int i, int state, int q) { st[(size_t)i * ss] = state; qv[(size_t)i * ss] = (uint8_t)q; };
    baq_glocal(A, br.xe - br.xb, L, br.bw, 0.001, 0.1, ref, qry, err, out);
    const bool extend = (flag >> 1) & 1;
    // BAM-orientation base i is the read's base (rev ? L-1-i : i); the output keeps the read's own orientation
    auto qual_at = [&](int i) { return (int)(rev ? ql[L - 1 - i] : ql[i]); };
    auto put = [&](int i, int bq) {
        const int q0 = qual_at(i);
        // qual -= (64 + (qual <= bq ? 0 : qual - bq)) - 64   (extended)   |   qual -= (qual - bq + 64) - 64   (plain; bq <= qual)
        qo[rev ? L - 1 - i : i] = (uint8_t)(extend ? (q0 <= bq ? q0 : bq) : bq);
    };
    int x = a.pos, y = 0;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, l = (int)(a.cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (!extend) {
                for (int i = y; i < y + l; ++i) {
                    const int s = st[(size_t)i * ss], q = qv[(size_t)i * ss], q0 = qual_at(i);
                    const int bq = ((s & 3) != 0 || s >> 2 != x - br.xb + (i - y)) ? 0 : (q0 < q ? q0 : q);
                    put(i, bq);
                }
            } else {
                // bq = agreement ? q : 0; left = running max from the block's start (kept in qv), right = from its end
                int run = 0;
                for (int i = y; i < y + l; ++i) {
                    const int s = st[(size_t)i * ss], q = qv[(size_t)i * ss];
                    const int bq = ((s & 3) != 0 || s >> 2 != x - br.xb + (i - y)) ? 0 : q;
                    st[(size_t)i * ss] = bq;
                    run = i == y ? bq : (bq > run ? bq : run);
                    qv[(size_t)i * ss] = (uint8_t)run;
                }
                for (int i = y + l - 1; i >= y; --i) {
                    const int bq = st[(size_t)i * ss];
                    run = i == y + l - 1 ? bq : (bq > run ? bq : run);
                    const int left = qv[(size_t)i * ss];
                    put(i, left < run ? left : run);
                }
            }
            x += l; y += l;
        } else if (op == 4 || op == 1) y += l;
        else if (op == 2) x += l;
    }
}

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
    baq_one(V, A, br, alns[br.read], codes, quals, stride, L, flag, quals_out, st + t, qv + t, (size_t)T);
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
    baq_one(V, A, br, alns[br.read], codes, quals, stride, L, flag, quals_out, st + (size_t)t * Lmax, qv + (size_t)t * Lmax, (size_t)1);
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
