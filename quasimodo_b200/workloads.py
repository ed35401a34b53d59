"""The five BASELINE.json configs as simulator settings.  Host-side only."""
import ctypes as C

import numpy as np

from . import _lib, genomes


def sim_params(seed, read_len=150, ins_mean=350, ins_sd=35, ins_max=700, indel_ppm=0, n_ppm=1000, lowq_ppm=20000):
    p = _lib.SimParams()
    p.seed, p.read_len, p.ins_mean, p.ins_sd, p.ins_max = seed, read_len, ins_mean, ins_sd, ins_max
    p.indel_ppm, p.n_ppm, p.lowq_ppm = indel_ppm, n_ppm, lowq_ppm
    return p


class Workload:
    """sources: list of (genome stem, copies); ref: list of genome stems concatenated into the index"""

    def __init__(self, name, sources, ref, n_pairs, seed, read_len=150, w=100, indel_ppm=0, ins_mean=350, ins_sd=35,
                 ins_max=700, extra_weights=None):
        self.name, self.ref_stems, self.n_pairs, self.w = name, ref, n_pairs, w
        self.sources = sources
        gs = [genomes.load(s) for s, _ in sources]
        self.src_codes = np.concatenate([g.codes for g in gs])
        lens = np.array([g.total for g in gs], dtype=np.int64)
        self.src_len = lens
        self.src_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        if extra_weights is not None:           # explicit read-share weights (e.g. contaminant spike-in)
            weights = [int(x) for x in extra_weights]
        else:                                   # strain chosen in proportion to copies x length
            weights = [int(c) * int(l) for (_, c), l in zip(sources, lens)]
        tot = sum(weights)
        acc, cum = 0, []
        for wgt in weights:
            acc += wgt
            cum.append(min((acc << 32) // tot, (1 << 32) - 1))
        cum[-1] = (1 << 32) - 1
        self.src_cum = np.array(cum, dtype=np.uint32)
        self.params = sim_params(seed, read_len, ins_mean, ins_sd, ins_max, indel_ppm)
        self.params.n_sources = len(sources)
        ref_g = genomes.load(ref[0])
        for s in ref[1:]:
            ref_g = ref_g.concat(genomes.load(s))
        self.ref = ref_g

    def simulate_host(self, pair0, n_pairs, stride=None):
        stride = stride or self.params.read_len
        codes = np.empty((2 * n_pairs, stride), dtype=np.uint8)
        quals = np.empty((2 * n_pairs, stride), dtype=np.uint8)
        src = np.empty(n_pairs, dtype=np.int32)
        pos = np.empty(n_pairs, dtype=np.int64)
        rc = _lib.lib().qm_simulate_pairs_host(C.byref(self.params), self.src_codes.ctypes.data, self.src_off.ctypes.data,
                                               self.src_len.ctypes.data, self.src_cum.ctypes.data, int(pair0), int(n_pairs),
                                               int(stride), codes.ctypes.data, quals.ctypes.data, src.ctypes.data, pos.ctypes.data)
        if rc:
            raise _lib.QmError(f"qm_simulate_pairs_host failed: {rc}")
        return codes, quals, src, pos


TA_SERIES = [("TA-1-0", 1, 0), ("TA-50-1", 50, 1), ("TA-10-1", 10, 1), ("TA-2-1", 2, 1), ("TA-1-1", 1, 1),
             ("TA-1-2", 1, 2), ("TA-1-10", 1, 10), ("TA-1-50", 1, 50), ("TA-1-100", 1, 100), ("TA-0-1", 0, 1)]


def config1(n_pairs=100_000):
    return Workload("cfg1:TM-1-1", [("TB40E", 1), ("Merlin", 1)], ["Merlin"], n_pairs, 1001)


def config2(i, n_pairs=2_000_000):
    name, t, a = TA_SERIES[i]
    src = [(s, c) for s, c in (("TB40E", t), ("AD169", a)) if c > 0]
    ref = "TB40E" if name == "TA-1-0" else "AD169"        # rules/load_config.smk:22-23
    return Workload("cfg2:" + name, src, [ref], n_pairs, 2000 + i)


def config3(n_pairs=1_000_000):
    # AD169:Merlin 1:10 + 5 % PhiX + 5 % E. coli read pairs; index = Merlin | phix | E. coli
    ad, me = genomes.load("AD169").total, genomes.load("Merlin").total
    hcmv = 1 * ad + 10 * me
    weights = [90 * 1 * ad, 90 * 10 * me, 5 * hcmv, 5 * hcmv]
    return Workload("cfg3:AM-1-10+phix+ecoli", [("AD169", 1), ("Merlin", 10), ("Phix", 1), ("Ecoli", 1)],
                    ["Merlin", "Phix", "Ecoli"], n_pairs, 3001, extra_weights=weights)


def config4(n_pairs=50_000_000):
    return Workload("cfg4:TM-1-50", [("TB40E", 1), ("Merlin", 50)], ["Merlin"], n_pairs, 4001)


def config5(n_pairs=2_000_000):
    return Workload("cfg5:MTA-10-3-1-2x250", [("Merlin", 10), ("TB40E", 3), ("AD169", 1)], ["Merlin"], n_pairs, 5001,
                    read_len=250, w=200, indel_ppm=200, ins_mean=550, ins_sd=50, ins_max=1000)
