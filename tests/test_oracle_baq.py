"""The BAQ restatement (oracle/qmo_baq.c: htslib realn.c sam_prob_realn over probaln.c kpa_glocal) checked on the
CPU against what the algorithm is known to do -- no htslib here, so these are properties, not golden vectors (parity unpinned):
the HMM recovers the true alignment of a clean read and of a read with a deletion, its posterior drops next to a mismatch and next
to an indel, BAQ never raises a quality, leaves bases outside M blocks alone, and extended BAQ >= plain BAQ."""
import numpy as np
import pytest

from oracle import qmo_py


def test_hmm_recovers_the_alignment_and_doubts_the_right_bases():
    rng = np.random.default_rng(1)
    ref = rng.integers(0, 4, 170).astype(np.uint8)
    ql = np.full(150, 35, np.uint8)
    q = ref[10:160].copy()
    st, bq, pr = qmo_py.kpa_glocal(ref, q, ql)
    assert np.array_equal(st >> 2, np.arange(10, 160)) and not (st & 3).any()
    assert bq[75] >= 60 and bq[0] < bq[75] and bq[-1] < bq[75]          # the ends could be shifted: lower posterior
    q2 = q.copy(); q2[70] = (q2[70] + 1) % 4
    st2, bq2, pr2 = qmo_py.kpa_glocal(ref, q2, ql)
    assert np.array_equal(st2 >> 2, np.arange(10, 160))
    assert bq2[70] < 40 <= bq[70] and pr2 > pr                          # the mismatch is doubted, the likelihood drops
    q3 = np.concatenate([ref[10:80], ref[83:163]])                       # 3 reference bases missing from the read
    st3, bq3, _ = qmo_py.kpa_glocal(ref, q3, ql)
    assert np.array_equal((st3 >> 2)[:60], np.arange(10, 70)) and np.array_equal((st3 >> 2)[80:], np.arange(93, 163))
    assert bq3[60:80].min() < 30                                         # where exactly the gap sits is uncertain
    # ambiguous bases match anything: the path does not move
    q4 = q.copy(); q4[40:43] = 4
    st4, _, _ = qmo_py.kpa_glocal(ref, q4, ql)
    assert np.array_equal(st4 >> 2, np.arange(10, 160))


@pytest.fixture(scope="module")
def case():
    from quasimodo_b200 import workloads
    n = 1500
    W = workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 250, np.int32)
    opt = qmo_py.default_opt()
    opt.w = 200
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, _, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)
    return dict(ref=ref, codes=codes, quals=quals, lens=lens, alns=alns, counts=counts)


def test_baq_caps_qualities_of_aligned_bases_only(case):
    c = case
    ext = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=3)
    plain = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=1)
    assert (ext <= c["quals"]).all() and (plain <= ext).all()
    assert (ext < c["quals"]).any() and (plain < ext).any()
    a = c["alns"]
    # unmapped / not admitted reads keep their qualities
    out = ((a["flag"] & 0x4) != 0) | (((a["flag"] & 0x1) != 0) & ((a["flag"] & 0x2) == 0))
    assert out.any() and np.array_equal(ext[out], c["quals"][out])
    # soft-clipped bases keep theirs (reads as sequenced: a forward read's leading clip is its first bases)
    r = next(i for i in range(len(a)) if not out[i] and a["n_cigar"][i] > 1 and (a["cigar"][i][0] & 0xf) == 4 and not (a["flag"][i] & 0x10))
    clip = int(a["cigar"][r][0] >> 4)
    assert np.array_equal(ext[r, :clip], c["quals"][r, :clip])
    # most bases of most reads are untouched at these qualities (BAQ bites at the ends and around indels)
    adm = ~out
    assert (ext[adm] == c["quals"][adm]).mean() > 0.8


def test_baq_changes_the_counts_only_downwards(case):
    c = case
    ext = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=3)
    with_baq = qmo_py.pileup(c["ref"], c["alns"], c["codes"], ext, c["lens"])
    base = c["counts"]
    assert np.array_equal(with_baq[:, 11:], base[:, 11:]) and np.array_equal(with_baq[:, 5], base[:, 5])   # depth / events: untouched
    # a lower quality can only drop a base below -Q ... except through the mate-overlap rule, which may now keep the other mate's base
    assert with_baq[:, :11].sum() < base[:, :11].sum()
