// simulate.cuh -- deterministic, index-addressable read-pair simulator (SURVEY.md 8d "Synthetic
// inputs").  The reference ships no reads (data/PRJEB32127.txt is a list of ENA URLs), so the
// benchmark inputs are simulated from the bundled genomes.  One function, compiled for host and
// device, integer arithmetic only => the same (seed, pair index) gives the same pair everywhere.
#pragma once
#include <stdint.h>
#include "../../include/quasimodo_b200.h"

#ifdef __CUDACC__
#define QM_HD __host__ __device__ __forceinline__
#else
#define QM_HD inline
#endif

// round(2^32 * 10^(-q/10)) for q = 0..41 (q=0 saturates)
QM_HD uint32_t qm_sim_err_threshold(int q)
{
    const uint32_t thr[42] = {
        4294967295u, 3411613790u, 2709941160u, 2152582778u, 1709857278u, 1358187913u, 1078847007u, 856958639u,
        680706443u, 540704347u, 429496730u, 341161379u, 270994116u, 215258278u, 170985728u, 135818791u,
        107884701u, 85695864u, 68070644u, 54070435u, 42949673u, 34116138u, 27099412u, 21525828u, 17098573u,
        13581879u, 10788470u, 8569586u, 6807064u, 5407043u, 4294967u, 3411614u, 2709941u, 2152583u, 1709857u,
        1358188u, 1078847u, 856959u, 680706u, 540704u, 429497u, 341161u};
    return thr[q < 0 ? 0 : (q > 41 ? 41 : q)];
}

struct qm_sim_rng {
    uint64_t s;
    QM_HD uint64_t next()
    {   // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

QM_HD int qm_sim_ctz16(uint32_t x)
{
    int n = 0;
    while (n < 16 && !((x >> n) & 1)) ++n;
    return n;
}

// walks the genome from `g` in direction dir (+1: forward strand as is, -1: reverse complement) and
// writes read_len bases / quals to out_b / out_q.
QM_HD void qm_sim_walk(const qm_sim_params &P, const uint8_t *genome, int64_t goff, int64_t glen, int64_t g, int dir,
                       qm_sim_rng &rng, uint8_t *out_b, uint8_t *out_q)
{
    const int L = P.read_len;
    int pending_ins = 0;
    for (int i = 0; i < L; ++i) {
        const uint64_t r = rng.next();
        int q = 38 - (8 * i) / (L > 1 ? L - 1 : 1);
        if (((r >> 34) & 0x3fff) < (uint64_t)((int64_t)P.lowq_ppm * 16384 / 1000000)) q = 2 + (int)((r >> 48) & 0xf) % 11;
        int b;
        if (P.indel_ppm > 0 && pending_ins == 0) {
            const uint64_t r2 = rng.next();
            if ((r2 & 0xfffff) < (uint64_t)((int64_t)P.indel_ppm * 1048576 / 1000000)) {
                int len = 1 + qm_sim_ctz16((uint32_t)(r2 >> 32) | 0x80u);
                if ((r2 >> 24) & 1) pending_ins = len;          // insertion: emit `len` random bases
                else g += (int64_t)dir * len;                   // deletion: skip genome bases
            }
        }
        if (pending_ins > 0) {
            b = (int)((r >> 32) & 3);
            --pending_ins;
        } else {
            if (g < 0 || g >= glen) b = 4;
            else {
                b = genome[goff + g];
                if (dir < 0) b = 3 - b;
            }
            g += dir;
            if (b < 4 && (uint32_t)r < qm_sim_err_threshold(q)) b = (b + 1 + (int)(((r >> 32) & 0xffff) % 3)) & 3;
        }
        if ((r >> 52) < (uint64_t)((int64_t)P.n_ppm * 4096 / 1000000)) b = 4;
        out_b[i] = (uint8_t)b;
        out_q[i] = (uint8_t)q;
    }
}

// pair `idx`: reads 2*idx (mate 1) and 2*idx+1 (mate 2); cum[] = cumulative 32-bit source thresholds
QM_HD void qm_sim_pair(const qm_sim_params &P, const uint8_t *genome, const int64_t *src_off, const int64_t *src_len,
                       const uint32_t *cum, int64_t idx, uint8_t *b1, uint8_t *q1, uint8_t *b2, uint8_t *q2,
                       int32_t *src_out, int64_t *pos_out)
{
    qm_sim_rng rng;
    rng.s = P.seed * 0xD6E8FEB86659FD93ull + (uint64_t)idx * 0xA24BAED4963EE407ull;
    const uint64_t r0 = rng.next(), r1 = rng.next(), r2 = rng.next();
    int s = 0;
    while (s + 1 < P.n_sources && (uint32_t)r0 >= cum[s]) ++s;
    const int64_t glen = src_len[s];
    const int64_t sum = (int64_t)(r1 & 0xffff) + (int64_t)((r1 >> 16) & 0xffff) + (int64_t)((r1 >> 32) & 0xffff) + (int64_t)((r1 >> 48) & 0xffff);
    int64_t ins = P.ins_mean + ((int64_t)P.ins_sd * (sum - 131070)) / 37837;
    if (ins < P.read_len) ins = P.read_len;
    if (ins > P.ins_max) ins = P.ins_max;
    if (ins > glen) ins = glen;
    const int64_t start = (int64_t)(r2 % (uint64_t)(glen - ins + 1));
    const int strand = (int)((r0 >> 32) & 1);
    // left walker: forward strand from start; right walker: reverse complement from start+ins-1
    if (strand == 0) {
        qm_sim_walk(P, genome, src_off[s], glen, start, +1, rng, b1, q1);
        qm_sim_walk(P, genome, src_off[s], glen, start + ins - 1, -1, rng, b2, q2);
    } else {
        qm_sim_walk(P, genome, src_off[s], glen, start + ins - 1, -1, rng, b1, q1);
        qm_sim_walk(P, genome, src_off[s], glen, start, +1, rng, b2, q2);
    }
    if (src_out) *src_out = s | (strand << 16);
    if (pos_out) *pos_out = start | (ins << 40);
}
