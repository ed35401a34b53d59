// index.cu -- k-mer hash index of the reference (replaces `bwa index`, rules/index.smk:13).
// One-off host build (the genomes are 0.24-4.9 Mb; SURVEY.md 8a1 "negligible"), device-resident tables:
// 16-byte open-addressing entries {k-mer, first, count} + ascending occurrence lists.  For the HCMV
// genomes the table is ~8 MB and stays in the 126 MB L2.
#include <algorithm>
#include <vector>
#include "pipeline.cuh"

extern "C" {

int qm_index_build(qm_ctx *ctx, const uint8_t *h_codes, int n_contigs, const int64_t *h_lens, int k, qm_index **out)
{
    if (!ctx || !out) return QM_EINVAL;
    *out = nullptr;
    if (!h_codes || !h_lens || n_contigs <= 0 || n_contigs > QM_MAX_CONTIGS || k < 8 || k > 31)
        return qm_fail(ctx, QM_EINVAL, "qm_index_build: need 1..%d contigs and 8 <= k <= 31 (the all-ones key marks an empty slot, which at k = 32 is the all-T k-mer)", QM_MAX_CONTIGS);
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    qm_index *ix = new qm_index();
    IndexView &v = ix->v;
    v.k = k; v.n_contigs = n_contigs;
    int64_t total = 0;
    for (int c = 0; c < n_contigs; ++c) {
        if (h_lens[c] <= 0) { delete ix; return qm_fail(ctx, QM_EINVAL, "contig %d has no bases", c); }
        v.off[c] = total; v.len[c] = h_lens[c]; total += h_lens[c];
    }
    if (total >= (1ll << 31)) { delete ix; return qm_fail(ctx, QM_ELIMIT, "reference too long for 32-bit positions"); }
    v.l_pac = total;
    for (int64_t i = 0; i < total; ++i)
        if (h_codes[i] > 3) { delete ix; return qm_fail(ctx, QM_EINVAL, "reference base %lld is not A/C/G/T", (long long)i); }

    // all k-mers that do not straddle a contig boundary
    std::vector<std::pair<uint64_t, uint32_t>> km;
    km.reserve((size_t)total);
    const uint64_t mask = k < 32 ? ((1ull << (2 * k)) - 1) : ~0ull;
    for (int c = 0; c < n_contigs; ++c) {
        uint64_t key = 0;
        for (int64_t p = 0; p < h_lens[c]; ++p) {
            key = ((key << 2) | h_codes[v.off[c] + p]) & mask;
            if (p >= k - 1) km.emplace_back(key, (uint32_t)(v.off[c] + p - (k - 1)));
        }
    }
    std::sort(km.begin(), km.end());
    ix->n_kmers = (int64_t)km.size();
    std::vector<uint32_t> pos(km.size() + 1);
    int64_t n_unique = 0;
    for (size_t i = 0; i < km.size(); ++i) {
        pos[i] = km[i].second;
        if (i == 0 || km[i].first != km[i - 1].first) ++n_unique;
    }
    ix->n_unique = n_unique;
    // load factor <= 0.25 while the table stays L2-sized (a miss then ends after ~1.4 probes instead of ~2.5), else <= 0.5
    int bits = 4;
    while ((1ll << bits) < 2 * n_unique) ++bits;
    if ((16ll << (bits + 1)) <= (64ll << 20)) ++bits;
    const int64_t tsize = 1ll << bits;
    ix->table_size = tsize;
    std::vector<uint4> table((size_t)tsize, make_uint4(0xffffffffu, 0xffffffffu, 0, 0));
    v.mask = (uint64_t)tsize - 1; v.shift = 64 - bits;
    for (size_t i = 0; i < km.size();) {
        size_t j = i;
        while (j < km.size() && km[j].first == km[i].first) ++j;
        const uint64_t key = km[i].first;
        uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> v.shift;
        while (table[h].x != 0xffffffffu || table[h].y != 0xffffffffu) h = (h + 1) & v.mask;
        // (a k-mer with ONE occurrence carries the position itself instead of its slot in pos[]: the seeding walk saves a dependent look-up)
        table[h] = make_uint4((uint32_t)key, (uint32_t)(key >> 32), (uint32_t)(j - i == 1 ? km[i].second : i), (uint32_t)(j - i));
        i = j;
    }
    // uniqueness bitmap: k-mer occurs once and its reverse complement never (see IndexView::uniq)
    std::vector<uint32_t> uniq((size_t)(total + 31) / 32 + 1, 0u), uniq2((size_t)(total + 31) / 32 + 1, 0u);
    std::vector<uint32_t> cnteqp[3];
    for (auto &m : cnteqp) m.assign((size_t)(total + 32 + 31) / 32 + 2, 0u);
    {
        auto count_of = [&](uint64_t key) -> uint32_t {
            uint64_t h = (key * 0x9E3779B97F4A7C15ull) >> v.shift;
            for (;;) {
                const uint4 e4 = table[h];
                const uint64_t kk = (uint64_t)e4.x | ((uint64_t)e4.y << 32);
                if (kk == key) return e4.w;
                if (e4.x == 0xffffffffu && e4.y == 0xffffffffu) return 0;
                h = (h + 1) & v.mask;
            }
        };
        for (size_t i = 0; i < km.size();) {
            size_t j = i;
            while (j < km.size() && km[j].first == km[i].first) ++j;
            {   // total occurrences on both strands in 2..4: the padded "count == T" bitmaps
                const uint64_t key = km[i].first;
                uint64_t rc = 0;
                for (int b = 0; b < k; ++b) rc |= (uint64_t)(3 - ((key >> (2 * b)) & 3)) << (2 * (k - 1 - b));
                if (rc != key) {
                    const size_t tot = (j - i) + count_of(rc);
                    if (tot >= 2 && tot <= 4)
                        for (size_t t = i; t < j; ++t) { const size_t x = (size_t)km[t].second + 32; cnteqp[tot - 2][x >> 5] |= 1u << (x & 31); }
                }
            }
            if (j - i == 1) {
                const uint64_t key = km[i].first;
                uint64_t rc = 0;
                for (int b = 0; b < k; ++b) rc |= (uint64_t)(3 - ((key >> (2 * b)) & 3)) << (2 * (k - 1 - b));
                // a palindromic k-mer (rc == key) is its own reverse complement: both look-ups hit, not unique
                if (rc != key) {
                    const uint32_t c_rc = count_of(rc);
                    if (c_rc == 0) uniq[km[i].second >> 5] |= 1u << (km[i].second & 31);
                    else if (c_rc == 1) uniq2[km[i].second >> 5] |= 1u << (km[i].second & 31);
                }
            }
            i = j;
        }
    }
    // padded 2-bit reference and padded uniqueness bitmap for the seeding kernel's 32-bases-at-a-time match extension
    std::vector<uint64_t> ref2p((size_t)(total + 32 + 31) / 32 + 2, 0ull);
    std::vector<uint32_t> uniqp((size_t)(total + 32 + 31) / 32 + 2, 0u), uniq2p((size_t)(total + 32 + 31) / 32 + 2, 0u);
    for (int64_t x = 0; x < total; ++x) {
        ref2p[(size_t)(x + 32) >> 5] |= (uint64_t)h_codes[x] << (2 * ((x + 32) & 31));
        if ((uniq[x >> 5] >> (x & 31)) & 1u) uniqp[(size_t)(x + 32) >> 5] |= 1u << ((x + 32) & 31);
        if ((uniq2[x >> 5] >> (x & 31)) & 1u) uniq2p[(size_t)(x + 32) >> 5] |= 1u << ((x + 32) & 31);
    }
    // Bloom filter over canonical k-mers: a read k-mer whose canonical form is not in the filter occurs on neither strand, and
    // the seeding walk answers that from three filter words instead of probing the table twice.  The filter lives in L2 (it
    // was sized for shared memory once: 208 KB whatever the genome, ~7 bits per k-mer of a herpesvirus and NO filter at all for the
    // 4.9 Mb index of config 3): 16 bits per distinct k-mer (3 hashes: ~0.5 % false positives), at least 208 KB, at most 256 MB.
    std::vector<uint32_t> bloom;
    v.bloom_bits = 0;
    const int64_t want_bits = std::max<int64_t>((int64_t)kBloomMinBytes * 8, ((n_unique * 16 + 8191) / 8192) * 8192);
    if (want_bits <= (int64_t)kBloomMaxBytes * 8) {
        const uint32_t bbits = (uint32_t)want_bits;
        bloom.assign(bbits / 32, 0u);
        for (size_t i = 0; i < km.size(); ++i) {
            if (i && km[i].first == km[i - 1].first) continue;
            const uint64_t key = km[i].first;
            uint64_t rc = 0;
            for (int b = 0; b < k; ++b) rc |= (uint64_t)(3 - ((key >> (2 * b)) & 3)) << (2 * (k - 1 - b));
            uint32_t p[3];
            qm_bloom_pos(key < rc ? key : rc, bbits, p);
            for (int t = 0; t < 3; ++t) bloom[p[t] >> 5] |= 1u << (p[t] & 31);
        }
        v.bloom_bits = bbits;
    }
    cudaError_t e;
    if (!bloom.empty() && ((e = cudaMalloc(&ix->d_bloom, bloom.size() * 4)) != cudaSuccess ||
                           (e = cudaMemcpy(ix->d_bloom, bloom.data(), bloom.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)) {
        qm_index_destroy(ctx, ix);
        return qm_fail(ctx, QM_ECUDA, "qm_index_build: %s", cudaGetErrorString(e));
    }
    v.bloom = (const uint32_t *)ix->d_bloom;
    if ((e = cudaMalloc(&ix->d_ref2p, ref2p.size() * 8)) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_ref2p, ref2p.data(), ref2p.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&ix->d_uniqp, uniqp.size() * 4)) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_uniqp, uniqp.data(), uniqp.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&ix->d_uniq2p, uniq2p.size() * 4)) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_uniq2p, uniq2p.data(), uniq2p.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        qm_index_destroy(ctx, ix);
        return qm_fail(ctx, QM_ECUDA, "qm_index_build: %s", cudaGetErrorString(e));
    }
    v.ref2p = (const uint64_t *)ix->d_ref2p; v.uniqp = (const uint32_t *)ix->d_uniqp; v.uniq2p = (const uint32_t *)ix->d_uniq2p;
    for (int t = 0; t < 3; ++t) {
        if ((e = cudaMalloc(&ix->d_cnteqp[t], cnteqp[t].size() * 4)) != cudaSuccess ||
            (e = cudaMemcpy(ix->d_cnteqp[t], cnteqp[t].data(), cnteqp[t].size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
            qm_index_destroy(ctx, ix);
            return qm_fail(ctx, QM_ECUDA, "qm_index_build: %s", cudaGetErrorString(e));
        }
        v.cnteqp[t] = (const uint32_t *)ix->d_cnteqp[t];
    }
    if ((e = cudaMalloc(&ix->d_uniq, uniq.size() * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_uniq, uniq.data(), uniq.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMalloc(&ix->d_refb, (size_t)total + 8)) != cudaSuccess ||        // + 8: word-wise readers (pileup.cu ld4) look one word past
        (e = cudaMemset(ix->d_refb, 0, (size_t)total + 8)) != cudaSuccess ||
        (e = cudaMalloc(&ix->d_table, (size_t)tsize * sizeof(uint4))) != cudaSuccess ||
        (e = cudaMalloc(&ix->d_pos, pos.size() * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_refb, h_codes, (size_t)total, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_table, table.data(), (size_t)tsize * sizeof(uint4), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(ix->d_pos, pos.data(), pos.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess) {
        qm_index_destroy(ctx, ix);
        return qm_fail(ctx, QM_ECUDA, "qm_index_build: %s", cudaGetErrorString(e));
    }
    v.refb = (const uint8_t *)ix->d_refb; v.table = (const uint4 *)ix->d_table; v.pos = (const uint32_t *)ix->d_pos;
    v.uniq = (const uint32_t *)ix->d_uniq;
    *out = ix;
    return QM_OK;
}

void qm_index_destroy(qm_ctx *ctx, qm_index *ix)
{
    if (!ix) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (ix->d_refb) cudaFree(ix->d_refb);
    if (ix->d_table) cudaFree(ix->d_table);
    if (ix->d_pos) cudaFree(ix->d_pos);
    if (ix->d_uniq) cudaFree(ix->d_uniq);
    if (ix->d_bloom) cudaFree(ix->d_bloom);
    if (ix->d_ref2p) cudaFree(ix->d_ref2p);
    if (ix->d_uniqp) cudaFree(ix->d_uniqp);
    if (ix->d_uniq2p) cudaFree(ix->d_uniq2p);
    for (int t = 0; t < 3; ++t) if (ix->d_cnteqp[t]) cudaFree(ix->d_cnteqp[t]);
    if (ix->d_fm_bwt) cudaFree(ix->d_fm_bwt);
    if (ix->d_fm_sa) cudaFree(ix->d_fm_sa);
    delete ix;
}

int64_t qm_index_lpac(const qm_index *ix) { return ix ? ix->v.l_pac : 0; }

}  // extern "C"
