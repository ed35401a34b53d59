"""GPU parity of base alignment quality (include/quasimodo_b200.h qm_baq_apply, qm_sample_set_baq) against the CPU
restatement of htslib's sam_prob_realn / kpa_glocal (oracle/qmo_baq.c): the capped qualities byte for byte -- double-precision
HMM, same operation order on both sides -- for plain and extended BAQ, reads with wide bands (a 15-base deletion) included, and the
sample's counts with BAQ on, immediate and deferred."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


def with_long_deletions(W, codes, n_edit):
    """pairs 0 .. n_edit-1 rewritten: read 1 = the reference with 15 bases cut out, read 2 = a clean mate 300 bases downstream"""
    ref = W.ref.codes
    L = codes.shape[1]
    rng = np.random.default_rng(7)
    for i in range(n_edit):
        p = int(rng.integers(1000, len(ref) // 2))
        codes[2 * i] = np.concatenate([ref[p:p + 120], ref[p + 135:p + 135 + L - 120]])
        mate = ref[p + 300:p + 300 + L]
        codes[2 * i + 1] = (3 - mate)[::-1]
    return codes


@pytest.fixture(scope="module", params=["cfg5", "cfg2"])
def case(request):
    from oracle import qmo_py
    from quasimodo_b200 import workloads
    if request.param == "cfg5":
        n, L = 4000, 250
        W = workloads.config5(n)
    else:
        n, L = 6000, 150
        W = workloads.config2(4, n)
    codes, quals, _, _ = W.simulate_host(0, n)
    if request.param == "cfg5":
        codes = with_long_deletions(W, codes.copy(), 40)
    lens = np.full(2 * n, L, np.int32)
    opt = qmo_py.default_opt()
    if request.param == "cfg5":
        opt.w = 200
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, _, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)
    return dict(name=request.param, W=W, n=n, codes=codes, quals=quals, lens=lens, alns=alns, ref=ref, w=opt.w)


def product_opt(case):
    from quasimodo_b200 import _lib
    opt = _lib.default_opt()
    opt.w = case["w"]
    return opt


@pytest.mark.parametrize("flag", [3, 1])
def test_baq_qualities_match_oracle(ctx, case, flag):
    from oracle import qmo_py
    c = case
    want = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=flag)
    idx = ctx.index(c["W"].ref, 31)
    got = ctx.baq_apply_host(idx, c["alns"], c["codes"], c["quals"], c["lens"], flag=flag)
    assert (want < c["quals"]).any()
    if c["name"] == "cfg5":
        # the wide-band class is exercised: records with a net indel above 7 bases
        a = c["alns"]
        net = np.zeros(len(a), np.int64)
        for k in range(a["cigar"].shape[1]):
            op, ln = a["cigar"][:, k] & 0xf, (a["cigar"][:, k] >> 4).astype(np.int64)
            live = k < a["n_cigar"]
            net += np.where(live & (op == 2), ln, 0) - np.where(live & (op == 1), ln, 0)
        wide = (np.abs(net) > 7) & ((a["flag"] & 0x2) != 0)
        assert wide.sum() >= 20
        assert np.array_equal(got[wide], want[wide])
    bad = np.argwhere(got != want)
    assert len(bad) == 0, (len(bad), bad[:5], got[tuple(bad[0])], want[tuple(bad[0])])
    idx.close()


def test_sample_counts_with_baq(ctx, case):
    from oracle import qmo_py
    c = case
    ext = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=3)
    want = qmo_py.pileup(c["ref"], c["alns"], c["codes"], ext, c["lens"])
    base = qmo_py.pileup(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"])
    assert not np.array_equal(want, base)
    idx = ctx.index(c["W"].ref, 31)
    s = ctx.sample(idx, product_opt(c))
    s.set_baq(3)
    s.add_pairs_host(c["codes"], c["quals"], c["lens"])
    assert np.array_equal(s.counts_host(), want)
    # deferred counting (a depth cap nothing reaches): the same counts through qm_sample_finish
    s.reset()
    s.set_max_depth(1 << 30)
    s.set_baq(3)
    s.add_pairs_host(c["codes"], c["quals"], c["lens"])
    s.finish()
    assert np.array_equal(s.counts_host(), want)
    # and off again: the -B counts
    s.reset()
    s.set_max_depth(0)
    s.set_baq(0)
    s.add_pairs_host(c["codes"], c["quals"], c["lens"])
    assert np.array_equal(s.counts_host(), base)
    s.close()
    idx.close()
