// fm_host.cpp -- HOST build of the device FM-index seeding (quasimodo_b200/csrc/fm_core.cuh) for the CPU test suite: the
// statements the kernel runs, driven read by read on an index given as bwa's .bwt / .sa bytes.  Test infrastructure.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../quasimodo_b200/csrc/fm_core.cuh"

struct HostCtg { int n; const int64_t *off, *len; int64_t l_pac; };

// bwt_bytes / sa_bytes: the two files as bwa writes them.  reads: n x stride base codes.  out: [n][max_seeds][3] int64
// (rbeg, qbeg, len), n_out[n].  Returns 0, or -1 on inconsistent index bytes.
extern "C" int fm_host_seeds(const uint8_t *bwt_bytes, int64_t bwt_len, const uint8_t *sa_bytes, int64_t sa_len, int n_contigs,
                             const int64_t *off, const int64_t *clen, int64_t l_pac, int min_seed_len, int max_occ, int max_mem_intv,
                             int64_t n, const uint8_t *reads, int stride, const int32_t *lens, int max_seeds, int64_t *out, int32_t *n_out)
{
    if (bwt_len < 40 || sa_len < 56) return -1;
    int64_t hdr[5], sh[7];
    memcpy(hdr, bwt_bytes, 40);
    memcpy(sh, sa_bytes, 56);
    FmView F;
    F.primary = hdr[0]; F.L2[0] = 0; memcpy(F.L2 + 1, hdr + 1, 32); F.seq_len = F.L2[4]; F.sa_intv = (int)sh[5];
    if (sh[0] != F.primary || sh[6] != F.seq_len || F.seq_len != 2 * l_pac) return -1;
    std::vector<uint32_t> bwt((size_t)(bwt_len - 40) / 4);
    memcpy(bwt.data(), bwt_bytes + 40, bwt.size() * 4);
    const int64_t n_sa = (F.seq_len + F.sa_intv) / F.sa_intv;
    std::vector<int64_t> sa((size_t)n_sa, -1);
    memcpy(sa.data() + 1, sa_bytes + 56, (size_t)(n_sa - 1) * 8);
    F.bwt = bwt.data(); F.sa = sa.data();
    HostCtg G = {n_contigs, off, clen, l_pac};
    std::vector<FmSeedOut> tmp((size_t)max_seeds);
    for (int64_t r = 0; r < n; ++r) {
        const int ns = fm_collect_seeds(F, G, min_seed_len, max_occ, max_mem_intv, lens[r], reads + r * stride, 1, tmp.data(), max_seeds);
        n_out[r] = ns;
        for (int i = 0; i < ns; ++i) { out[(r * max_seeds + i) * 3] = tmp[i].rbeg; out[(r * max_seeds + i) * 3 + 1] = tmp[i].qbeg; out[(r * max_seeds + i) * 3 + 2] = tmp[i].len; }
    }
    return 0;
}
