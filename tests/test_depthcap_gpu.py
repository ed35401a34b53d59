"""GPU parity of the depth cap (`bcftools mpileup -d N`): the replay of htslib's pileup iterator in the
library (csrc/depthcap.cu) against the oracle's restatement (oracle/qmo_pileup.c qmo_depth_cap) on a deep sample of a small
genome -- the same reads dropped, the same counts -- alone and combined with duplicate removal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def deep():
    from oracle import qmo_py
    from quasimodo_b200 import workloads
    n = 40_000                                      # 2200x on PhiX
    W = workloads.Workload("phix-deep", [("Phix", 1)], ["Phix"], n, 77)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    codes[1:2400:2] = 4                             # unplaced second mates: orphans are not admitted
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, _, _ = qmo_py.run_sample(ref, codes, quals, lens)
    return dict(W=W, n=n, codes=codes, quals=quals, lens=lens, ref=ref, alns=alns, counts=counts)


@pytest.mark.parametrize("max_depth", [250, 1000, 8000])
def test_depth_cap_matches_oracle(ctx, deep, max_depth):
    from oracle import qmo_py
    want_counts, keep = qmo_py.pileup_capped(deep["ref"], deep["alns"], deep["codes"], deep["quals"], deep["lens"], max_depth)
    admitted = qmo_py.depth_cap(deep["ref"], deep["alns"], 1 << 30)
    idx = ctx.index(deep["W"].ref, 31)
    s = ctx.sample(idx)
    s.set_max_depth(max_depth)
    h = deep["n"] // 3 * 2                          # two chunks: the cap runs over the whole sample
    s.add_pairs_host(deep["codes"][:h], deep["quals"][:h], deep["lens"][:h])
    s.add_pairs_host(deep["codes"][h:], deep["quals"][h:], deep["lens"][h:], pair_id0=h // 2)
    assert not s.counts_host().any()
    n_dup, n_capped = s.finish()
    assert n_dup == 0 and n_capped == int((admitted & ~keep).sum())
    if max_depth < 8000:
        assert 0 < n_capped < admitted.sum()
    else:
        assert n_capped == 0 and np.array_equal(want_counts, deep["counts"])
    assert np.array_equal(s.counts_host(), want_counts)
    s.close()
    idx.close()


def test_depth_cap_after_duplicate_removal(ctx, deep):
    """the reference's order: picard removes duplicates (rule rmdup), bcftools mpileup then caps what is left"""
    from oracle import dedup_py, qmo_py
    alns = deep["alns"]
    dup = dedup_py.mark_duplicates(alns, deep["quals"], deep["lens"])
    marked = alns.copy()
    for e in (0, 1):
        sel = dup & ((marked["flag"][e::2] & 4) == 0)
        marked["flag"][e::2][sel] |= 0x400
    want_counts, keep = qmo_py.pileup_capped(deep["ref"], marked, deep["codes"], deep["quals"], deep["lens"], 500)
    idx = ctx.index(deep["W"].ref, 31)
    s = ctx.sample(idx)
    s.set_rmdup(True)
    s.set_max_depth(500)
    s.add_pairs_host(deep["codes"], deep["quals"], deep["lens"])
    n_dup, n_capped = s.finish()
    assert n_dup == int(dup.sum()) > 0 and n_capped > 0
    assert np.array_equal(s.counts_host(), want_counts)
    s.close()
    idx.close()
