// baq_host.cpp -- HOST build of the base-alignment-quality logic (quasimodo_b200/csrc/baq_core.cuh) for the CPU test suite: the
// statements the CUDA kernels baq_fast_kernel / baq_slow_kernel run, over a bounds-checked scratch.  Test infrastructure (built by
// tests/test_baq_host.py into tests/_build/); the product has no host path.
#include <stdint.h>
#include <math.h>
#include <vector>
#include "../quasimodo_b200/csrc/baq_core.cuh"

namespace {
// the slab layout of baq.cu's SlowAcc (rows of the read's own width), every access checked
struct HostAcc {
    std::vector<double> *v;
    int row, L;
    double &f(int i, int c) const { return v->at((size_t)i * row + c); }
    double &b(int j, int c) const { return v->at((size_t)(L + 1 + j) * row + c); }
    double &s(int i) const { return v->at((size_t)(L + 3) * row + i); }
};
}

// quals_out = quals with the admitted reads' qualities capped, as qm_baq_apply does it; returns the number of reads in the wide-band class
extern "C" long long baq_host(const uint8_t *ref, const int64_t *ctg_off, const int64_t *ctg_len, const qm_pileup_opt *po, const qm_aln *alns,
                              const uint8_t *codes, const uint8_t *quals, int stride, const int32_t *lens, long long n_reads, int flag,
                              uint8_t *quals_out)
{
    float q2p[256];
    for (int i = 0; i < 256; ++i) q2p[i] = (float)pow(10, -i / 10.);
    for (long long i = 0; i < n_reads * stride; ++i) quals_out[i] = quals[i];
    long long wide = 0;
    std::vector<double> work;
    std::vector<int32_t> st((size_t)stride);
    std::vector<uint8_t> qv((size_t)stride);
    for (long long r = 0; r < n_reads; ++r) {
        const qm_aln &a = alns[r];
        if (!baq_admitted(*po, a)) continue;
        const int L = lens[r];
        int xb, xe, bw;
        if (!baq_window(a, L, ctg_len[a.rid], xb, xe, bw)) continue;
        const int hb = baq_band(xe - xb, L, bw);
        wide += hb != 7;
        const int row = (2 * hb + 1) * 3 + 6;
        work.assign((size_t)(L + 3) * row + (L + 2), -1.0);          // (poisoned: the logic must initialise what it reads)
        HostAcc A = {&work, row, L};
        const BaqRead br = {(int32_t)r, xb, xe, bw};
        baq_one(ref + ctg_off[a.rid] + xb, q2p, A, br, a, codes, quals, stride, L, flag, quals_out, st.data(), qv.data(), (size_t)1);
    }
    return wide;
}
