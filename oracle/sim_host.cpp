// sim_host.cpp -- the read-pair simulator for the CPU arms (test infrastructure, like the rest of oracle/).
// The simulator is the benchmark's INPUT generator (the reference ships no reads), one header compiled for host and
// device (quasimodo_b200/csrc/simulate.cuh); this file builds its host half into oracle/libqmsim.so so that the CPU arm
// of bench.py (--impl reference) generates its inputs without mapping the product library.
#include "../quasimodo_b200/csrc/simulate.cuh"

extern "C" int qmsim_pairs_host(const qm_sim_params *p, const uint8_t *genome, const int64_t *src_off, const int64_t *src_len,
                                const uint32_t *src_cum, int64_t pair0, int64_t n_pairs, int32_t stride, uint8_t *codes,
                                uint8_t *quals, int32_t *src, int64_t *pos)
{
    if (!p || !genome || !src_off || !src_len || !src_cum || !codes || !quals || n_pairs < 0 || stride < p->read_len || p->n_sources < 1)
        return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_pairs; ++i) {
        uint8_t *b1 = codes + (2 * i) * stride, *b2 = b1 + stride;
        uint8_t *q1 = quals + (2 * i) * stride, *q2 = q1 + stride;
        qm_sim_pair(*p, genome, src_off, src_len, src_cum, pair0 + i, b1, q1, b2, q2, src ? src + i : nullptr, pos ? pos + i : nullptr);
        for (int j = p->read_len; j < stride; ++j) { b1[j] = 4; b2[j] = 4; q1[j] = 0; q2[j] = 0; }
    }
    return 0;
}
