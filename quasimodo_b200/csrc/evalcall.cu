// evalcall.cu -- (1) SNP calls from the count tensor, (2) the TP/FP/FN matcher.
//
// (1) stands in for `bcftools call -p 0.01 --ploidy 1 -mv | bcftools view -i 'INFO/DP>=10'`
//     (rules/vcfcall.smk:116-117).  bcftools' multiallelic likelihood model and its QUAL are NOT reproduced
//     (outside the parity contract); this is a threshold caller whose records carry what the
//     downstream consumers read: POS, single-base REF/ALT, QUAL, DP, AF.
// (2) replaces the three bash pipelines of program/extract_TP_FP_SNPs.py:24-57 -- `fgrep -wf` of
//     "POS\t.\tREF\tALT" patterns against the caller's SNP lines -- by a sorted-key membership test:
//     key = pos << 8 | ref << 4 | alt (CHROM is not compared, exactly like the script; B.5).
#include <math.h>
#include <iterator>
#include <cub/device/device_scan.cuh>
#include "pipeline.cuh"

namespace {

__device__ __forceinline__ bool call_test(const qm_call_opt &o, int dp_raw, int tot, int ad)
{
    return dp_raw >= o.min_dp && ad >= o.min_alt && tot > 0 && (double)ad >= (double)o.min_af * (double)tot;
}

// pass 0: number of calls per position; pass 1: write them at offs[pos]
template <int PASS>
__global__ void call_kernel(IndexView V, qm_call_opt o, const int32_t *__restrict__ counts, int *__restrict__ n_at,
                            const int64_t *__restrict__ offs, qm_call *__restrict__ out, int64_t max_calls)
{
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= V.l_pac) return;
    const int64_t L = V.l_pac;
    int ad[4], adf[4], tot = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) { adf[b] = counts[b * L + p]; ad[b] = adf[b] + counts[(6 + b) * L + p]; tot += ad[b]; }
    const int dp_raw = counts[14 * L + p];
    const int ref = V.refb[p];
    int n = 0;
    int64_t at = PASS ? offs[p] : 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if (b == ref || !call_test(o, dp_raw, tot, ad[b])) continue;
        if (PASS) {
            if (at < max_calls) {
                qm_call c;
                const int rid = qm_pos2rid(V, p);
                c.rid = rid; c.pos = (int32_t)(p - V.off[rid]);
                c.ref = (uint8_t)ref; c.alt = (uint8_t)b; c.pad[0] = c.pad[1] = 0;
                c.dp = dp_raw;
                c.ad_ref_f = adf[ref]; c.ad_ref_r = ad[ref] - adf[ref];
                c.ad_alt_f = adf[b]; c.ad_alt_r = ad[b] - adf[b];
                // QUAL: Chernoff bound on the binomial tail P(X >= ad | tot, e = 0.002), phred scaled, capped at 999
                const double f = (double)ad[b] / tot, e = 0.002;
                double kl = f * log(f / e);
                if (f < 1.0) kl += (1.0 - f) * log((1.0 - f) / (1.0 - e));
                double q = f > e ? 4.342944819032518 * tot * kl : 0.0;
                c.qual = (float)(q > 999.0 ? 999.0 : q);
                c.af = (float)f;
                out[at] = c;
            }
            ++at;
        }
        ++n;
    }
    if (!PASS) n_at[p] = n;
}

// widening iterator for the device-wide scan below: int counts read as int64
struct WidenInt {
    const int *p;
    using value_type = int64_t; using difference_type = int64_t; using pointer = const int64_t *; using reference = int64_t;
    using iterator_category = std::random_access_iterator_tag;
    __host__ __device__ int64_t operator[](int64_t i) const { return i < n ? (int64_t)p[i] : 0; }
    __host__ __device__ int64_t operator*() const { return (*this)[0]; }
    __host__ __device__ WidenInt operator+(int64_t k) const { WidenInt r = *this; r.p += k; r.n -= k; return r; }
    int64_t n;
};

// ---- matcher ----
__global__ void pad_copy_kernel(const uint64_t *__restrict__ in, int64_t n, int64_t n_pad, uint64_t *__restrict__ out)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = i < n ? in[i] : ~0ull;
}

// one block sorts n_pad (power of two) keys in global memory: bitonic network, __syncthreads between passes
__global__ void __launch_bounds__(1024) bitonic_kernel(uint64_t *__restrict__ a, int64_t n_pad)
{
    for (int64_t k = 2; k <= n_pad; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            for (int64_t t = threadIdx.x; t < (n_pad >> 1); t += blockDim.x) {
                const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const bool up = (lo & k) == 0;
                const uint64_t x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
}

__global__ void member_kernel(const uint64_t *__restrict__ q, int64_t nq, const uint64_t *__restrict__ sorted, int64_t ns,
                              uint8_t *__restrict__ flags)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint64_t key = q[i];
    int64_t lo = 0, hi = ns;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted[mid] < key) lo = mid + 1; else hi = mid; }
    flags[i] = (lo < ns && sorted[lo] == key) ? 1 : 0;
}

// qm_call records -> matcher keys (1-based POS << 8 | ref << 4 | alt)
__global__ void call_keys_kernel(const qm_call *__restrict__ calls, int64_t n, uint64_t *__restrict__ keys)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const qm_call c = calls[i];
    keys[i] = ((uint64_t)(uint32_t)(c.pos + 1) << 8) | ((uint64_t)c.ref << 4) | (uint64_t)c.alt;
}

// out[which] += number of set flags
__global__ void __launch_bounds__(256) flag_count_kernel(const uint8_t *__restrict__ flags, int64_t n, unsigned long long *__restrict__ out)
{
    unsigned v = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v += flags[i] != 0;
    v = __reduce_add_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, (unsigned long long)v);
}

int sorted_copy(qm_ctx *ctx, int which, const uint64_t *d_keys, int64_t n, uint64_t **out, cudaStream_t st)
{
    int64_t n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, which, (size_t)n_pad * 8, &p);
    if (rc) return rc;
    pad_copy_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(d_keys, n, n_pad, (uint64_t *)p);
    bitonic_kernel<<<1, 1024, 0, st>>>((uint64_t *)p, n_pad);
    *out = (uint64_t *)p;
    return QM_OK;
}

}  // namespace

extern "C" {

void qm_call_opt_default(qm_call_opt *o) { o->min_dp = 10; o->min_alt = 2; o->min_af = 0.01f; o->reserved = 0; }

int qm_call_snps(qm_ctx *ctx, const qm_index *idx, const qm_call_opt *copt, const int32_t *d_counts, qm_call *d_calls,
                 int64_t max_calls, int64_t *h_n_calls, void *stream)
{
    if (!ctx || !idx || !copt || !d_counts || !h_n_calls || max_calls < 0 || (max_calls > 0 && !d_calls)) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t L = idx->v.l_pac;
    void *p = nullptr;
    const size_t n_bytes = ((size_t)L * 4 + 255) & ~(size_t)255;
    // exclusive scan of the per-column call counts over L + 1 items (the last input reads as 0): offs[L] = the total.
    // Device-wide (one block took 4 ms on the 4.9 Mb index of config 3).
    size_t cub_bytes = 0;
    WidenInt in_probe = {nullptr, 0};
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, in_probe, (int64_t *)nullptr, (int)(L + 1), st);
    const size_t o_cub = (n_bytes + (size_t)(L + 1) * 8 + 255) & ~(size_t)255;
    int rc = qm_scratch_reserve(ctx, 7, o_cub + cub_bytes, &p);
    if (rc) return rc;
    int *n_at = (int *)p;
    int64_t *offs = (int64_t *)((char *)p + n_bytes);
    const unsigned grid = (unsigned)((L + 127) / 128);
    const int sp = qm_prof_begin(ctx, QM_ST_OTHER, st);
    call_kernel<0><<<grid, 128, 0, st>>>(idx->v, *copt, d_counts, n_at, nullptr, nullptr, 0);
    { WidenInt in = {n_at, L}; QM_CUDA(ctx, cub::DeviceScan::ExclusiveSum((char *)p + o_cub, cub_bytes, in, offs, (int)(L + 1), st)); }
    call_kernel<1><<<grid, 128, 0, st>>>(idx->v, *copt, d_counts, nullptr, offs, d_calls, max_calls);
    qm_prof_end(ctx, QM_ST_OTHER, sp, st, 3);
    QM_CUDA(ctx, cudaGetLastError());
    QM_CUDA(ctx, cudaMemcpyAsync(h_n_calls, offs + L, 8, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    if (*h_n_calls > max_calls) return qm_fail(ctx, QM_ELIMIT, "qm_call_snps: %lld calls exceed max_calls=%lld", (long long)*h_n_calls, (long long)max_calls);
    return QM_OK;
}

int qm_eval_match(qm_ctx *ctx, const uint64_t *d_call_keys, int64_t n_call, const uint64_t *d_truth_keys, int64_t n_truth,
                  uint8_t *d_call_flags, uint8_t *d_truth_flags, void *stream)
{
    if (!ctx || n_call < 0 || n_truth < 0 || (n_call > 0 && (!d_call_keys || !d_call_flags)) || (n_truth > 0 && !d_truth_keys))
        return QM_EINVAL;
    if (n_call > (1ll << 26) || n_truth > (1ll << 26)) return qm_fail(ctx, QM_ELIMIT, "qm_eval_match: more than 2^26 keys");
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int sp = qm_prof_begin(ctx, QM_ST_OTHER, st);
    int launches = 0;
    if (n_call > 0) {
        if (n_truth > 0) {
            uint64_t *sorted = nullptr;
            int rc = sorted_copy(ctx, 8, d_truth_keys, n_truth, &sorted, st);
            if (rc) return rc;
            member_kernel<<<(unsigned)((n_call + 255) / 256), 256, 0, st>>>(d_call_keys, n_call, sorted, n_truth, d_call_flags);
            launches += 3;
        } else QM_CUDA(ctx, cudaMemsetAsync(d_call_flags, 0, (size_t)n_call, st));
    }
    if (n_truth > 0 && d_truth_flags) {
        if (n_call > 0) {
            uint64_t *sorted = nullptr;
            int rc = sorted_copy(ctx, 9, d_call_keys, n_call, &sorted, st);
            if (rc) return rc;
            member_kernel<<<(unsigned)((n_truth + 255) / 256), 256, 0, st>>>(d_truth_keys, n_truth, sorted, n_call, d_truth_flags);
            launches += 3;
        } else QM_CUDA(ctx, cudaMemsetAsync(d_truth_flags, 0, (size_t)n_truth, st));
    }
    qm_prof_end(ctx, QM_ST_OTHER, sp, st, launches);
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

// calls of one sample against the truth keys, everything on the device: keys from the qm_call records, both membership
// passes, the three totals.  h_tp_fp_fn = {TP, FP, FN}.  d_call_flags (may be NULL) receives the per-call TP flags.  Synchronous.
int qm_eval_calls(qm_ctx *ctx, const qm_call *d_calls, int64_t n_call, const uint64_t *d_truth_keys, int64_t n_truth,
                  uint8_t *d_call_flags, int64_t h_tp_fp_fn[3], void *stream)
{
    if (!ctx || !h_tp_fp_fn || n_call < 0 || n_truth < 0 || (n_call > 0 && !d_calls) || (n_truth > 0 && !d_truth_keys)) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_keys = 0, o_cf = al((size_t)n_call * 8), o_tf = o_cf + al((size_t)n_call), o_cnt = o_tf + al((size_t)n_truth);
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 19, o_cnt + 256, &p);
    if (rc) return rc;
    char *b = (char *)p;
    uint64_t *keys = (uint64_t *)(b + o_keys);
    uint8_t *cf = d_call_flags ? d_call_flags : (uint8_t *)(b + o_cf), *tf = (uint8_t *)(b + o_tf);
    unsigned long long *cnt = (unsigned long long *)(b + o_cnt);
    QM_CUDA(ctx, cudaMemsetAsync(cnt, 0, 16, st));
    if (n_call) call_keys_kernel<<<(unsigned)((n_call + 255) / 256), 256, 0, st>>>(d_calls, n_call, keys);
    rc = qm_eval_match(ctx, keys, n_call, d_truth_keys, n_truth, cf, tf, st);
    if (rc) return rc;
    if (n_call) flag_count_kernel<<<(unsigned)((n_call + 4095) / 4096 < 1024 ? (n_call + 4095) / 4096 : 1024), 256, 0, st>>>(cf, n_call, cnt);
    if (n_truth) flag_count_kernel<<<(unsigned)((n_truth + 4095) / 4096 < 1024 ? (n_truth + 4095) / 4096 : 1024), 256, 0, st>>>(tf, n_truth, cnt + 1);
    QM_CUDA(ctx, cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    QM_CUDA(ctx, cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    h_tp_fp_fn[0] = (int64_t)h[0]; h_tp_fp_fn[1] = n_call - (int64_t)h[0]; h_tp_fp_fn[2] = n_truth - (int64_t)h[1];
    return QM_OK;
}

// same with host buffers: copies keys in, flags out; synchronous
int qm_eval_match_host(qm_ctx *ctx, const uint64_t *h_call_keys, int64_t n_call, const uint64_t *h_truth_keys, int64_t n_truth,
                       uint8_t *h_call_flags, uint8_t *h_truth_flags)
{
    if (!ctx || n_call < 0 || n_truth < 0 || (n_call > 0 && (!h_call_keys || !h_call_flags)) || (n_truth > 0 && !h_truth_keys))
        return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t kc = ((size_t)n_call * 8 + 255) & ~(size_t)255, kt = ((size_t)n_truth * 8 + 255) & ~(size_t)255;
    const size_t fc = ((size_t)n_call + 255) & ~(size_t)255, ft = ((size_t)n_truth + 255) & ~(size_t)255;
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 11, kc + kt + fc + ft + 256, &p);
    if (rc) return rc;
    char *b = (char *)p;
    uint64_t *dc = (uint64_t *)b, *dt = (uint64_t *)(b + kc);
    uint8_t *dfc = (uint8_t *)(b + kc + kt), *dft = (uint8_t *)(b + kc + kt + fc);
    cudaStream_t st = ctx->own_stream;
    if (n_call) QM_CUDA(ctx, cudaMemcpyAsync(dc, h_call_keys, (size_t)n_call * 8, cudaMemcpyHostToDevice, st));
    if (n_truth) QM_CUDA(ctx, cudaMemcpyAsync(dt, h_truth_keys, (size_t)n_truth * 8, cudaMemcpyHostToDevice, st));
    rc = qm_eval_match(ctx, dc, n_call, dt, n_truth, dfc, h_truth_flags ? dft : nullptr, st);
    if (rc) return rc;
    if (n_call) QM_CUDA(ctx, cudaMemcpyAsync(h_call_flags, dfc, (size_t)n_call, cudaMemcpyDeviceToHost, st));
    if (n_truth && h_truth_flags) QM_CUDA(ctx, cudaMemcpyAsync(h_truth_flags, dft, (size_t)n_truth, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    return QM_OK;
}

}  // extern "C"
