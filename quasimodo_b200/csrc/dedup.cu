// dedup.cu -- duplicate marking with the rule of `picard MarkDuplicates` (the reference's rule `rmdup`,
// rules/rmdup.smk:13-16, which sits between `bwa` and every caller; semantics SURVEY.md B.9).
//
// A pair with both ends placed is keyed by {contig, unclipped 5' coordinate, strand} of both ends (ends in
// coordinate order); among pairs with the same key the one with the highest sum of base qualities >= 15 over both
// mates stays, ties go to the earliest pair of the input.  A read whose mate is unplaced ("fragment") is keyed by its
// own {contig, unclipped 5' coordinate, strand}: it is a duplicate whenever an end of some fully placed pair has that
// key, otherwise the best fragment of the key stays.  Pairs with no placed end are never duplicates.
// On the device: one 64-bit key + score per pair, the stable radix sort of sort.cu, then one pass over the sorted
// runs.  Duplicates get SAM flag 0x400, which the pileup's read admission already skips.
#include "pipeline.cuh"

namespace {

constexpr uint64_t kNoKey = ~0ull;
constexpr int kCoordBias = 4096;                 // unclipped coordinates can be negative (clipped bases before the contig)

// {contig, unclipped 5' coordinate} of a placed record as a 30-bit code, and its strand
// (26 bits of coordinate, 4 of contig: a contig of >= 2^26 - 4096 bases or a 17th contig does not fit; *bad is raised then
// and qm_mark_duplicates fails with QM_ELIMIT instead of letting keys collide)
__device__ __forceinline__ uint32_t end_code(const qm_aln &a, int *rev, int *bad)
{
    const bool r = (a.flag & 0x10) != 0;
    int lead = 0, trail = 0, rlen = 0;
    const int nc = a.n_cigar == 255 ? 0 : a.n_cigar;
    for (int k = 0; k < nc; ++k) {
        const int op = a.cigar[k] & 0xf, len = (int)(a.cigar[k] >> 4);
        if (op == 0 || op == 2) rlen += len;
    }
    if (nc > 0 && (a.cigar[0] & 0xf) == 4) lead = (int)(a.cigar[0] >> 4);
    if (nc > 1 && (a.cigar[nc - 1] & 0xf) == 4) trail = (int)(a.cigar[nc - 1] >> 4);
    const int coord = r ? a.pos + rlen - 1 + trail : a.pos - lead;
    *rev = r ? 1 : 0;
    if (coord + kCoordBias < 0 || coord + kCoordBias >= (1 << 26) || a.rid >= 16) *bad = 1;
    return ((uint32_t)a.rid << 26) | (uint32_t)(coord + kCoordBias);
}

__device__ __forceinline__ bool placed(const qm_aln &a) { return !(a.flag & 0x4) && a.rid >= 0 && a.n_cigar != 0 && a.n_cigar != 255; }

// key, score and (for fully placed pairs) the two end keys of every pair
__global__ void __launch_bounds__(128)
dup_keys_kernel(const qm_aln *__restrict__ alns, const uint8_t *__restrict__ quals, int stride, const int32_t *__restrict__ lens,
                int64_t n_pairs, int64_t pair0, uint64_t *__restrict__ keys, int32_t *__restrict__ scores, uint64_t *__restrict__ ends,
                unsigned long long *__restrict__ n_bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const qm_aln a = alns[2 * i], b = alns[2 * i + 1];
    const bool pa = placed(a), pb = placed(b);
    int sc[2] = {0, 0};
    for (int e = 0; e < 2; ++e) {
        const uint8_t *q = quals + (2 * i + e) * (int64_t)stride;
        const int L = lens[2 * i + e];
        int s = 0;
        for (int j = 0; j < L; ++j) { const int v = q[j]; s += v >= 15 ? v : 0; }
        sc[e] = s;
    }
    uint64_t key = kNoKey, e0 = kNoKey, e1 = kNoKey;
    int score = 0, bad = 0;
    if (pa && pb) {
        int ra, rb;
        uint32_t ca = end_code(a, &ra, &bad), cb = end_code(b, &rb, &bad);
        if (cb < ca || (cb == ca && rb < ra)) { const uint32_t t = ca; ca = cb; cb = t; const int u = ra; ra = rb; rb = u; }
        key = ((uint64_t)ca << 32) | ((uint64_t)cb << 2) | (uint64_t)(ra << 1 | rb);
        e0 = (uint64_t)ca << 1 | (uint64_t)ra; e1 = (uint64_t)cb << 1 | (uint64_t)rb;
        score = sc[0] + sc[1];
    } else if (pa || pb) {
        int r;
        const uint32_t c = end_code(pa ? a : b, &r, &bad);
        key = (1ull << 63) | ((uint64_t)c << 1) | (uint64_t)r;
        score = pa ? sc[0] : sc[1];
    }
    if (bad) atomicAdd(n_bad, 1ull);
    keys[pair0 + i] = key;
    scores[pair0 + i] = score;
    ends[2 * (pair0 + i)] = e0; ends[2 * (pair0 + i) + 1] = e1;
}

// one thread per sorted position: the head of a run of equal keys picks the run's representative (highest score, first in
// input order on ties -- the sort is stable, so that is the first of the run) and flags the others
__global__ void __launch_bounds__(128)
dup_mark_kernel(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ perm, const int32_t *__restrict__ scores, int64_t n,
                const uint64_t *__restrict__ sorted_ends, int64_t n_ends, uint8_t *__restrict__ dup)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = skeys[i];
    if (k == kNoKey) return;
    if (i > 0 && skeys[i - 1] == k) return;             // not the head of its run
    if (k >> 63) {                                      // fragments: any end of a fully placed pair with this key wins
        const uint64_t want = k & ~(1ull << 63);
        int64_t lo = 0, hi = n_ends;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (sorted_ends[mid] < want) lo = mid + 1; else hi = mid; }
        if (lo < n_ends && sorted_ends[lo] == want) {
            for (int64_t j = i; j < n && skeys[j] == k; ++j) dup[perm[j]] = 1;
            return;
        }
    }
    int64_t best = i;
    int bs = scores[perm[i]];
    for (int64_t j = i + 1; j < n && skeys[j] == k; ++j) { const int s = scores[perm[j]]; if (s > bs) { bs = s; best = j; } }
    for (int64_t j = i; j < n && skeys[j] == k; ++j) if (j != best) dup[perm[j]] = 1;
}

__global__ void __launch_bounds__(128)
dup_apply_kernel(qm_aln *__restrict__ alns, int64_t n_pairs, int64_t pair0, const uint8_t *__restrict__ dup, unsigned long long *__restrict__ n_dup)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs || !dup[pair0 + i]) return;
    for (int e = 0; e < 2; ++e)
        if (!(alns[2 * i + e].flag & 0x4)) alns[2 * i + e].flag |= 0x400;      // an unplaced mate is never a duplicate
    atomicAdd(n_dup, 1ull);
}

}  // namespace

extern "C" {

// chunks: the sample's records in input order, chunk c = pairs [pair0[c], pair0[c] + n[c]).  Synchronous.
int qm_mark_duplicates(qm_ctx *ctx, int n_chunks, qm_aln *const *d_alns, const uint8_t *const *d_quals, const int32_t *strides,
                       const int32_t *const *d_lens, const int64_t *n_pairs, int64_t *h_n_dup, void *stream)
{
    if (!ctx || n_chunks < 0 || (n_chunks > 0 && (!d_alns || !d_quals || !strides || !d_lens || !n_pairs))) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t N = 0;
    for (int c = 0; c < n_chunks; ++c) { if (n_pairs[c] < 0) return QM_EINVAL; N += n_pairs[c]; }
    if (h_n_dup) *h_n_dup = 0;
    if (N == 0) return QM_OK;
    if (2 * N > 0xffffffffll) return qm_fail(ctx, QM_ELIMIT, "qm_mark_duplicates: more than 2^31 pairs");
    // scratch 15: keys | ends | scores | perm | end perm | dup flags | counter
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_keys = 0, o_ends = o_keys + al((size_t)N * 8), o_sc = o_ends + al((size_t)2 * N * 8), o_perm = o_sc + al((size_t)N * 4);
    const size_t o_eperm = o_perm + al((size_t)N * 4), o_dup = o_eperm + al((size_t)2 * N * 4), o_cnt = o_dup + al((size_t)N);
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 15, o_cnt + 256, &p);
    if (rc) return rc;
    char *b = (char *)p;
    uint64_t *keys = (uint64_t *)(b + o_keys), *ends = (uint64_t *)(b + o_ends);
    int32_t *scores = (int32_t *)(b + o_sc);
    uint32_t *perm = (uint32_t *)(b + o_perm), *eperm = (uint32_t *)(b + o_eperm);
    uint8_t *dup = (uint8_t *)(b + o_dup);
    unsigned long long *cnt = (unsigned long long *)(b + o_cnt);
    QM_CUDA(ctx, cudaMemsetAsync(dup, 0, (size_t)N, st));
    QM_CUDA(ctx, cudaMemsetAsync(cnt, 0, 16, st));             // cnt[0] = duplicates, cnt[1] = pairs whose end code does not fit
    int64_t p0 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        if (n_pairs[c]) dup_keys_kernel<<<(unsigned)((n_pairs[c] + 127) / 128), 128, 0, st>>>(d_alns[c], d_quals[c], strides[c], d_lens[c], n_pairs[c], p0,
                                                                                        keys, scores, ends, cnt + 1);
        p0 += n_pairs[c];
    }
    rc = qm_sort_pairs(ctx, keys, perm, N, 64, st);
    if (rc) return rc;
    rc = qm_sort_pairs(ctx, ends, eperm, 2 * N, 32, st);          // end keys are 31 bits; kNoKey's low word is all ones: sorts last
    if (rc) return rc;
    dup_mark_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(keys, perm, scores, N, ends, 2 * N, dup);
    p0 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        if (n_pairs[c]) dup_apply_kernel<<<(unsigned)((n_pairs[c] + 127) / 128), 128, 0, st>>>(d_alns[c], n_pairs[c], p0, dup, cnt);
        p0 += n_pairs[c];
    }
    QM_CUDA(ctx, cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    QM_CUDA(ctx, cudaMemcpyAsync(h, cnt, 16, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    if (h[1]) return qm_fail(ctx, QM_ELIMIT, "qm_mark_duplicates: %llu pairs lie beyond what the 30-bit end code holds (contigs of up to 2^26 - 4096 bases, 16 contigs)", h[1]);
    if (h_n_dup) *h_n_dup = (int64_t)h[0];
    return QM_OK;
}

}  // extern "C"
