// fm_core.cuh -- bwa-mem's own seeding through bwa's own index: FM-index bidirectional search on the device.
// (bwa 0.7.17 bwt.c: bwt_occ4 / bwt_2occ4 / bwt_extend / bwt_smem1a / bwt_seed_strategy1 / bwt_sa, bwamem.c:
// mem_collect_intv and the seed loop of mem_chain; reference call site rules/bwa.smk:15, index files ref/*.bwt + ref/*.sa
// written by rules/index.smk:13)
//
// The default seeder of this library is a k-mer hash index: all exact matches >= k, a superset of bwa's SMEM
// positions that differs from bwa's seed set in repeats and re-seeding corner cases.  This file is the alternative that makes
// the seeds bwa's: the three rounds of mem_collect_intv (super-maximal exact matches; long rare SMEMs searched again from
// their middle for more frequent matches; the LAST-like round that takes the shortest match of >= min_seed_len bases with
// fewer than max_mem_intv occurrences), intervals sorted by their span on the read, every interval turned into seeds by
// walking the sampled suffix array, in the order mem_chain visits them.  The index is bwa's: either the reference's own
// ref/*.bwt + ref/*.sa bytes, or the same bytes rebuilt from the genome (index.cu; checked equal in the tests).
//
// One THREAD per read: the search is a chain of dependent occurrence-count look-ups (two 32-byte checkpoints + up to 127
// packed bases each) into a table of a few hundred kB that lives in L2; reads of a batch are independent, so the kernel
// hides the latency with many reads in flight rather than with parallelism inside one search.  Plain C++ over an index
// view, compiled for the device by fmseed.cu and for the host by tests/fm_host.cpp (the CPU suite checks it against the
// oracle's restatement seed for seed).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FM_HD __host__ __device__ __forceinline__
#else
#define FM_HD inline
#endif

struct FmView {
    const uint32_t *bwt;          // per 128 rows: four int64 occurrence counts (8 words) then 8 words of 16 two-bit symbols
    const int64_t *sa;            // sa[r / sa_intv] = text position of matrix row r (r a multiple of sa_intv); sa[0] unused
    int64_t primary, seq_len;     // row of the sentinel; 2 x l_pac
    int64_t L2[5];                // cumulative symbol counts
    int sa_intv;
};

struct FmIv { uint32_t k, l, s; uint32_t info; };       // bi-interval [k, k+s) / [l, l+s); info = read span begin << 16 | end

constexpr int kFmMaxStack = 48;                         // interval sizes one forward search can pass through
constexpr int kFmMaxIv = 96;                            // intervals collected per read over the three rounds

FM_HD int fm_sym(const FmView &F, int64_t r) { return (int)(F.bwt[(r >> 7 << 4) + 8 + ((r & 0x7f) >> 4)] >> ((~r & 0xf) << 1) & 3); }

// symbol counts in matrix rows [0, r] (r = -1: none); the sentinel row is not stored, rows behind it shift by one
FM_HD void fm_occ4(const FmView &F, int64_t r, uint32_t cnt[4])
{
    if (r < 0) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    r -= (r >= F.primary);
    const uint32_t *blk = F.bwt + (r >> 7 << 4);
    cnt[0] = blk[0]; cnt[1] = blk[2]; cnt[2] = blk[4]; cnt[3] = blk[6];          // low words of the four int64 counts
    const int rem = (int)(r & 0x7f);                                             // symbols [0, rem] of the block
    for (int w = 0; w <= rem >> 4; ++w) {
        const uint32_t x = blk[8 + w];
        const int n = w < rem >> 4 ? 16 : (rem & 15) + 1;                        // leading symbols of this word that count (MSB first)
        // per-symbol counts of a word of 2-bit symbols: one marker bit per symbol, compared against each pattern, popcount
        const uint32_t lo = x & 0x55555555u, hi = (x >> 1) & 0x55555555u;
        const uint32_t live = n == 16 ? 0x55555555u : 0x55555555u & ~(0x55555555u >> (n << 1));
        const uint32_t is3 = hi & lo & live, is2 = hi & ~lo & live, is1 = ~hi & lo & live, is0 = ~hi & ~lo & live;
#if defined(__CUDA_ARCH__)
        cnt[0] += __popc(is0); cnt[1] += __popc(is1); cnt[2] += __popc(is2); cnt[3] += __popc(is3);
#else
        cnt[0] += (uint32_t)__builtin_popcount(is0); cnt[1] += (uint32_t)__builtin_popcount(is1);
        cnt[2] += (uint32_t)__builtin_popcount(is2); cnt[3] += (uint32_t)__builtin_popcount(is3);
#endif
    }
}

FM_HD FmIv fm_single(const FmView &F, int c)
{
    FmIv v;
    v.k = (uint32_t)F.L2[c] + 1; v.s = (uint32_t)(F.L2[c + 1] - F.L2[c]); v.l = (uint32_t)F.L2[3 - c] + 1; v.info = 0;
    return v;
}

// the four one-symbol extensions of a bi-interval: backward (prepend a read base) or forward (append; the caller passes the
// complement and the roles of k and l swap)
FM_HD void fm_extend(const FmView &F, const FmIv &in, FmIv out[4], bool back)
{
    const uint32_t a = back ? in.k : in.l, b = back ? in.l : in.k;
    uint32_t lo[4], hi[4];
    fm_occ4(F, (int64_t)a - 1, lo);
    fm_occ4(F, (int64_t)a - 1 + in.s, hi);
    uint32_t run = b + (uint32_t)((int64_t)a <= F.primary && (int64_t)a + in.s - 1 >= F.primary);
    for (int c = 3; c >= 0; --c) {
        const uint32_t na = (uint32_t)F.L2[c] + 1 + lo[c], ns = hi[c] - lo[c];
        out[c].s = ns; out[c].info = in.info;
        if (back) { out[c].k = na; out[c].l = run; } else { out[c].l = na; out[c].k = run; }
        run += ns;
    }
}

// Super-maximal exact matches through read position x (bwt_smem1a, max_intv = 0), appended to mem[] by start; returns the
// position the next search starts at.  min_s: smallest interval size to keep extending (1 in round one, occurrences + 1 in
// round two).
FM_HD int fm_smem(const FmView &F, int len, const uint8_t *q, int qstep, int x, uint32_t min_s, FmIv *mem, int &n_mem, int max_mem)
{
    FmIv stack_a[kFmMaxStack], stack_b[kFmMaxStack];
    FmIv *prev = stack_a, *curr = stack_b;
    int n_prev = 0, n_curr = 0;
    const int first = n_mem;
    auto base = [&](int i) { const int c = q[(int64_t)i * qstep]; return c > 3 ? 4 : c; };
    if (base(x) > 3) return x + 1;
    if (min_s < 1) min_s = 1;
    FmIv cur = fm_single(F, base(x)), nxt[4];
    cur.info = (uint32_t)(x + 1);
    int i;
    for (i = x + 1; i < len; ++i) {                    // forward: remember the match every time its interval shrinks
        const int c = base(i);
        if (c > 3) { if (n_curr < kFmMaxStack) curr[n_curr++] = cur; break; }
        fm_extend(F, cur, nxt, false);
        if (nxt[3 - c].s != cur.s) {
            if (n_curr < kFmMaxStack) curr[n_curr++] = cur;
            if (nxt[3 - c].s < min_s) break;
        }
        cur = nxt[3 - c]; cur.info = (uint32_t)(i + 1);
    }
    if (i == len && n_curr < kFmMaxStack) curr[n_curr++] = cur;
    for (int a = 0, b = n_curr - 1; a < b; ++a, --b) { const FmIv t = curr[a]; curr[a] = curr[b]; curr[b] = t; }      // longest match first
    const int ret = (int)curr[0].info;
    { FmIv *t = curr; curr = prev; prev = t; n_prev = n_curr; }
    for (i = x - 1; i >= -1; --i) {                    // backward: a match that cannot grow and is not inside a kept one is an SMEM
        const int c = i < 0 ? 4 : base(i);
        n_curr = 0;
        for (int j = 0; j < n_prev; ++j) {
            const FmIv p = prev[j];
            if (c <= 3) fm_extend(F, p, nxt, true);
            if (c > 3 || nxt[c].s < min_s) {
                if (n_curr == 0 && (n_mem == first || i + 1 < (int)(mem[n_mem - 1].info >> 16))) {
                    FmIv m = p;
                    m.info = (uint32_t)(i + 1) << 16 | (p.info & 0xffffu);
                    if (n_mem < max_mem) mem[n_mem++] = m;
                }
            } else if (n_curr == 0 || nxt[c].s != curr[n_curr - 1].s) {
                FmIv e = nxt[c];
                e.info = p.info;
                curr[n_curr++] = e;
            }
        }
        if (n_curr == 0) break;
        { FmIv *t = curr; curr = prev; prev = t; n_prev = n_curr; }
    }
    for (int a = first, b = n_mem - 1; a < b; ++a, --b) { const FmIv t = mem[a]; mem[a] = mem[b]; mem[b] = t; }        // by start on the read
    return ret;
}

// text position of matrix row r: walk backwards through the text until a sampled row (bwt_sa / bwt_invPsi)
FM_HD int64_t fm_locate(const FmView &F, int64_t r)
{
    int64_t steps = 0;
    const int64_t mask = F.sa_intv - 1;
    while (r & mask) {
        ++steps;
        if (r == F.primary) { r = 0; continue; }
        const int c = fm_sym(F, r - (r > F.primary));
        uint32_t cnt[4];
        fm_occ4(F, r, cnt);
        r = F.L2[c] + cnt[c];
    }
    return steps + F.sa[r / F.sa_intv];
}

struct FmSeedOut { int64_t rbeg; int32_t qbeg, len; };

// All three rounds for one read, then the seeds in mem_chain's order (intervals by read span; inside an interval every
// step-th row so that at most max_occ positions are taken; a seed that bridges two contigs or the strand boundary is not one).
// Ctg: contig table with n, off[], len[] (forward coordinates), l_pac.  Returns the number of seeds written (<= max_seeds).
template <class Ctg>
FM_HD int fm_collect_seeds(const FmView &F, const Ctg &G, int min_seed_len, int max_occ, int max_mem_intv, int len, const uint8_t *q, int qstep,
                           FmSeedOut *out, int max_seeds)
{
    FmIv mem[kFmMaxIv];
    int n = 0;
    auto base = [&](int i) { const int c = q[(int64_t)i * qstep]; return c > 3 ? 4 : c; };
    // round one: SMEMs of the whole read
    for (int x = 0; x < len;) {
        if (base(x) > 3) { ++x; continue; }
        const int from = n;
        x = fm_smem(F, len, q, qstep, x, 1, mem, n, kFmMaxIv);
        int w = from;
        for (int i = from; i < n; ++i) if ((int)(mem[i].info & 0xffffu) - (int)(mem[i].info >> 16) >= min_seed_len) mem[w++] = mem[i];
        n = w;
    }
    // round two: a long SMEM with few occurrences, searched again from its middle for matches that occur more often
    const int split_len = (int)(min_seed_len * 1.5f + .499f), n1 = n;
    for (int k = 0; k < n1; ++k) {
        const int b = (int)(mem[k].info >> 16), e = (int)(mem[k].info & 0xffffu);
        if (e - b < split_len || mem[k].s > 10) continue;
        const int from = n;
        fm_smem(F, len, q, qstep, (b + e) >> 1, mem[k].s + 1, mem, n, kFmMaxIv);
        int w = from;
        for (int i = from; i < n; ++i) if ((int)(mem[i].info & 0xffffu) - (int)(mem[i].info >> 16) >= min_seed_len) mem[w++] = mem[i];
        n = w;
    }
    // round three: from every position the shortest match of at least min_seed_len bases with fewer than max_mem_intv hits
    if (max_mem_intv > 0)
        for (int x = 0; x < len;) {
            if (base(x) > 3) { ++x; continue; }
            FmIv cur = fm_single(F, base(x)), nxt[4];
            int i, next = len;
            bool found = false;
            for (i = x + 1; i < len; ++i) {
                const int c = base(i);
                if (c > 3) { next = i + 1; break; }
                fm_extend(F, cur, nxt, false);
                if (nxt[3 - c].s < (uint32_t)max_mem_intv && i - x >= min_seed_len) {
                    cur = nxt[3 - c]; cur.info = (uint32_t)x << 16 | (uint32_t)(i + 1);
                    found = true; next = i + 1;
                    break;
                }
                cur = nxt[3 - c];
            }
            if (found && cur.s > 0 && n < kFmMaxIv) mem[n++] = cur;
            x = next;
        }
    // order of bwa's introsort on (start << 32 | end): insertion sort, spans are unique enough that stability does not matter
    // beyond equal keys keeping their round order
    for (int i = 1; i < n; ++i) {
        const FmIv v = mem[i];
        int j = i - 1;
        while (j >= 0 && mem[j].info > v.info) { mem[j + 1] = mem[j]; --j; }
        mem[j + 1] = v;
    }
    int ns = 0;
    for (int i = 0; i < n && ns < max_seeds; ++i) {
        const int qb = (int)(mem[i].info >> 16), sl = (int)(mem[i].info & 0xffffu) - qb;
        const uint32_t step = mem[i].s > (uint32_t)max_occ ? mem[i].s / (uint32_t)max_occ : 1u;
        uint32_t kk = 0;
        for (int count = 0; kk < mem[i].s && count < max_occ && ns < max_seeds; kk += step, ++count) {
            const int64_t rb = fm_locate(F, (int64_t)mem[i].k + kk), re = rb + sl;
            if (rb < G.l_pac && re > G.l_pac) continue;                                  // across the strand boundary
            const int64_t fb = rb >= G.l_pac ? 2 * G.l_pac - re : rb, fe = fb + sl;
            bool inside = false;
            for (int c = 0; c < G.n; ++c) inside |= fb >= G.off[c] && fe <= G.off[c] + G.len[c];
            if (!inside) continue;                                                       // across two contigs
            out[ns].rbeg = rb; out[ns].qbeg = qb; out[ns].len = sl;
            ++ns;
        }
    }
    return ns;
}
