// fmindex.cu -- bwa's FM-index as a device-resident part of a qm_index (the alternative seeder of fm_core.cuh).
// Two ways in: qm_index_attach_bwa takes the bytes of the reference's own index files (ref/X.bwt and ref/X.sa, written by
// `bwa index`, rules/index.smk:13) -- the drop-in case, a Snakemake deployment already has them next to every genome; and
// qm_index_build_fm rebuilds the same bytes from the genome (suffix array of forward + reverse complement by prefix doubling,
// BWT without the sentinel row, occurrence checkpoints every 128 rows interleaved with the packed symbols, suffix-array samples
// every 32 rows), which the tests compare with digests of the reference's files.  Host work, once per genome.
#include <string.h>
#include <algorithm>
#include <numeric>
#include <vector>
#include "pipeline.cuh"

namespace {

// suffix array of t[0, n): ranks by the first 2^r symbols, refined by doubling; only ties are re-sorted
std::vector<int32_t> suffix_sort(const std::vector<uint8_t> &t)
{
    const int32_t n = (int32_t)t.size();
    std::vector<int32_t> sa((size_t)n), rk((size_t)n), tmp((size_t)n);
    std::iota(sa.begin(), sa.end(), 0);
    auto key8 = [&](int32_t i) { uint32_t v = 0; for (int j = 0; j < 8; ++j) v = v * 5 + (i + j < n ? t[(size_t)i + j] + 1u : 0u); return v; };
    std::vector<uint32_t> k8((size_t)n);
    for (int32_t i = 0; i < n; ++i) k8[(size_t)i] = key8(i);
    std::sort(sa.begin(), sa.end(), [&](int32_t a, int32_t b) { return k8[(size_t)a] < k8[(size_t)b]; });
    rk[(size_t)sa[0]] = 0;
    for (int32_t i = 1; i < n; ++i) rk[(size_t)sa[(size_t)i]] = rk[(size_t)sa[(size_t)i - 1]] + (k8[(size_t)sa[(size_t)i]] != k8[(size_t)sa[(size_t)i - 1]]);
    for (int64_t h = 8; h < n; h <<= 1) {
        auto second = [&](int32_t i) { return (int64_t)i + h < n ? rk[(size_t)(i + h)] : -1; };
        bool any = false;
        for (int32_t b = 0; b < n;) {
            int32_t e = b + 1;
            while (e < n && rk[(size_t)sa[(size_t)e]] == rk[(size_t)sa[(size_t)b]]) ++e;
            if (e - b > 1) {
                std::sort(sa.begin() + b, sa.begin() + e, [&](int32_t x, int32_t y) { return second(x) < second(y); });
                any = true;
            }
            b = e;
        }
        if (!any) break;
        tmp[(size_t)sa[0]] = 0;
        for (int32_t i = 1; i < n; ++i) {
            const int32_t x = sa[(size_t)i - 1], y = sa[(size_t)i];
            tmp[(size_t)y] = tmp[(size_t)x] + (rk[(size_t)x] != rk[(size_t)y] || second(x) != second(y));
        }
        rk.swap(tmp);
        if (rk[(size_t)sa[(size_t)n - 1]] == n - 1) break;
    }
    return sa;
}

}  // namespace

extern "C" {

// h_bwt / h_sa: the two files as bwa 0.7.17 writes them (.bwt: int64 primary, int64 L2[1..4], then per 128 rows four int64
// occurrence counts + eight uint32 words of sixteen 2-bit symbols; .sa: int64 primary, four int64, int64 sa_intv, int64 seq_len,
// then the samples of rows sa_intv, 2 sa_intv, ...).  The index must be of the same genome: seq_len = 2 x l_pac is checked.
int qm_index_attach_bwa(qm_ctx *ctx, qm_index *idx, const uint8_t *h_bwt, int64_t bwt_bytes, const uint8_t *h_sa, int64_t sa_bytes)
{
    if (!ctx || !idx || !h_bwt || !h_sa) return QM_EINVAL;
    if (bwt_bytes < 40 || sa_bytes < 56) return qm_fail(ctx, QM_EINVAL, "qm_index_attach_bwa: truncated index files");
    int64_t hdr[5], sh[7];
    memcpy(hdr, h_bwt, 40);
    memcpy(sh, h_sa, 56);
    FmView F;
    F.primary = hdr[0]; F.L2[0] = 0; memcpy(F.L2 + 1, hdr + 1, 32); F.seq_len = F.L2[4]; F.sa_intv = (int)sh[5];
    if (sh[0] != F.primary || sh[6] != F.seq_len) return qm_fail(ctx, QM_EINVAL, "qm_index_attach_bwa: the .bwt and the .sa are not of one index");
    if (F.seq_len != 2 * idx->v.l_pac)
        return qm_fail(ctx, QM_EINVAL, "qm_index_attach_bwa: the index covers %lld symbols, this genome needs %lld (forward + reverse complement)",
                       (long long)F.seq_len, (long long)(2 * idx->v.l_pac));
    if (F.sa_intv < 1 || (F.sa_intv & (F.sa_intv - 1))) return qm_fail(ctx, QM_EINVAL, "qm_index_attach_bwa: suffix-array interval %d is not a power of two", F.sa_intv);
    const int64_t n_words = (bwt_bytes - 40) / 4, need_words = ((F.seq_len + 127) / 128 + 1) * 8 + (F.seq_len + 15) / 16;
    const int64_t n_sa = (F.seq_len + F.sa_intv) / F.sa_intv;
    if (n_words < need_words - 8 || sa_bytes < 56 + 8 * (n_sa - 1)) return qm_fail(ctx, QM_EINVAL, "qm_index_attach_bwa: index files shorter than their headers say");
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (idx->d_fm_bwt) { cudaFree(idx->d_fm_bwt); idx->d_fm_bwt = nullptr; }
    if (idx->d_fm_sa) { cudaFree(idx->d_fm_sa); idx->d_fm_sa = nullptr; }
    std::vector<int64_t> sa((size_t)n_sa, -1);
    memcpy(sa.data() + 1, h_sa + 56, (size_t)(n_sa - 1) * 8);
    QM_CUDA(ctx, cudaMalloc(&idx->d_fm_bwt, (size_t)n_words * 4 + 64));
    QM_CUDA(ctx, cudaMemset(idx->d_fm_bwt, 0, (size_t)n_words * 4 + 64));
    QM_CUDA(ctx, cudaMemcpy(idx->d_fm_bwt, h_bwt + 40, (size_t)n_words * 4, cudaMemcpyHostToDevice));
    QM_CUDA(ctx, cudaMalloc(&idx->d_fm_sa, (size_t)n_sa * 8));
    QM_CUDA(ctx, cudaMemcpy(idx->d_fm_sa, sa.data(), (size_t)n_sa * 8, cudaMemcpyHostToDevice));
    F.bwt = (const uint32_t *)idx->d_fm_bwt; F.sa = (const int64_t *)idx->d_fm_sa;
    idx->fm = F;
    idx->have_fm = true;
    idx->fm_bwt_bytes.assign(h_bwt, h_bwt + 40 + n_words * 4);
    idx->fm_sa_bytes.assign(h_sa, h_sa + 56 + 8 * (n_sa - 1));
    return QM_OK;
}

// the same index built here from the genome (h_codes as given to qm_index_build)
int qm_index_build_fm(qm_ctx *ctx, qm_index *idx, const uint8_t *h_codes)
{
    if (!ctx || !idx || !h_codes) return QM_EINVAL;
    const int64_t l_pac = idx->v.l_pac, n = 2 * l_pac;
    if (n >= (1ll << 31)) return qm_fail(ctx, QM_ELIMIT, "qm_index_build_fm: genome too long for 32-bit suffix positions");
    std::vector<uint8_t> t((size_t)n);
    for (int64_t i = 0; i < l_pac; ++i) { t[(size_t)i] = h_codes[i]; t[(size_t)(n - 1 - i)] = (uint8_t)(3 - h_codes[i]); }
    const std::vector<int32_t> sa = suffix_sort(t);
    // matrix rows: row 0 is the sentinel suffix, row r > 0 suffix sa[r - 1]; the row of suffix 0 (what precedes it is the
    // sentinel) is `primary` and is not stored
    std::vector<uint8_t> sym((size_t)n);
    int64_t primary = 0, w = 0, L2[5] = {0, 0, 0, 0, 0};
    for (int64_t r = 0; r <= n; ++r) {
        const int64_t sfx = r == 0 ? n : sa[(size_t)r - 1];
        if (sfx == 0) { primary = r; continue; }
        sym[(size_t)w++] = t[(size_t)sfx - 1];
    }
    for (int64_t i = 0; i < n; ++i) ++L2[sym[(size_t)i] + 1];
    for (int c = 0; c < 4; ++c) L2[c + 1] += L2[c];
    std::vector<uint8_t> bwt_file(40), sa_file(56);
    { const int64_t hdr[5] = {primary, L2[1], L2[2], L2[3], L2[4]}; memcpy(bwt_file.data(), hdr, 40); }
    std::vector<uint32_t> words;
    int64_t cnt[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n; ++i) {
        if ((i & 127) == 0) { const uint32_t *c32 = (const uint32_t *)cnt; words.insert(words.end(), c32, c32 + 8); }
        if ((i & 15) == 0) words.push_back(0u);
        words.back() |= (uint32_t)sym[(size_t)i] << ((~i & 15) << 1);
        ++cnt[sym[(size_t)i]];
    }
    { const uint32_t *c32 = (const uint32_t *)cnt; words.insert(words.end(), c32, c32 + 8); }
    bwt_file.resize(40 + words.size() * 4);
    memcpy(bwt_file.data() + 40, words.data(), words.size() * 4);
    const int intv = 32;
    { const int64_t sh[7] = {primary, L2[1], L2[2], L2[3], L2[4], intv, n}; memcpy(sa_file.data(), sh, 56); }
    for (int64_t r = intv; r <= n; r += intv) {
        const int64_t v = sa[(size_t)r - 1];
        const uint8_t *b = (const uint8_t *)&v;
        sa_file.insert(sa_file.end(), b, b + 8);
    }
    return qm_index_attach_bwa(ctx, idx, bwt_file.data(), (int64_t)bwt_file.size(), sa_file.data(), (int64_t)sa_file.size());
}

// the attached index as the two files bwa would write (sizes with NULL buffers)
int qm_index_fm_export(const qm_index *idx, uint8_t *h_bwt, int64_t *bwt_bytes, uint8_t *h_sa, int64_t *sa_bytes)
{
    if (!idx || !bwt_bytes || !sa_bytes) return QM_EINVAL;
    if (!idx->have_fm) return QM_EINVAL;
    *bwt_bytes = (int64_t)idx->fm_bwt_bytes.size(); *sa_bytes = (int64_t)idx->fm_sa_bytes.size();
    if (h_bwt) memcpy(h_bwt, idx->fm_bwt_bytes.data(), idx->fm_bwt_bytes.size());
    if (h_sa) memcpy(h_sa, idx->fm_sa_bytes.data(), idx->fm_sa_bytes.size());
    return QM_OK;
}

}  // extern "C"
