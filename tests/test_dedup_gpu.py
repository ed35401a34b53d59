"""GPU parity of duplicate marking (picard MarkDuplicates REMOVE_DUPLICATES=true, rules/rmdup.smk:13-16) against the
CPU restatement, on a deep sample of a small genome where a few percent of the pairs are positional duplicates; then the counts
of the survivors against the oracle pileup, and the driver's --rmdup-bam."""
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
    """PhiX (5.4 kb) at ~1700x: 30,000 pairs, plus 600 pairs whose second mate is blanked to N (fragments)"""
    from oracle import dedup_py, qmo_py
    from quasimodo_b200 import workloads
    n = 30_000
    W = workloads.Workload("phix-deep", [("Phix", 1)], ["Phix"], n, 901)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    codes[1:1200:2] = 4                                     # 600 unplaced second mates
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts_all, _, _ = qmo_py.run_sample(ref, codes, quals, lens)
    dup = dedup_py.mark_duplicates(alns, quals, lens)
    marked = alns.copy()
    for e in (0, 1):
        sel = dup & ((marked["flag"][e::2] & 4) == 0)
        marked["flag"][e::2][sel] |= 0x400
    counts = qmo_py.pileup(ref, marked, codes, quals, lens)
    return dict(W=W, n=n, codes=codes, quals=quals, lens=lens, alns=alns, dup=dup, marked=marked, counts=counts, counts_all=counts_all)


def test_rmdup_matches_oracle(ctx, deep):
    assert 0.02 < deep["dup"].mean() < 0.6 and deep["dup"][:600].any()
    idx = ctx.index(deep["W"].ref, 31)
    s = ctx.sample(idx)
    s.set_rmdup(True)
    half = deep["n"] // 2 * 2                                # two chunks: duplicates are found across them
    s.add_pairs_host(deep["codes"][:half], deep["quals"][:half], deep["lens"][:half])
    s.add_pairs_host(deep["codes"][half:], deep["quals"][half:], deep["lens"][half:], pair_id0=half // 2)
    assert not s.counts_host().any()                         # nothing is counted before the duplicates are known
    n_dup = s.rmdup_finish()
    got = s.kept_alns(deep["n"])
    assert n_dup == int(deep["dup"].sum())
    assert np.array_equal(got["flag"], deep["marked"]["flag"])
    for f in ("rid", "pos", "mapq", "n_cigar", "nm", "tlen"):
        assert np.array_equal(got[f], deep["alns"][f]), f
    assert np.array_equal(s.counts_host(), deep["counts"])
    assert deep["counts"][:, 14].sum() < deep["counts_all"][:, 14].sum()
    # a finished sample takes neither a second finish nor more pairs until it is reset (its kept chunks are gone)
    from quasimodo_b200 import QmError
    with pytest.raises(QmError):
        s.rmdup_finish()
    with pytest.raises(QmError):
        s.add_pairs_host(deep["codes"][:4], deep["quals"][:4], deep["lens"][:4])
    assert np.array_equal(s.counts_host(), deep["counts"])
    s.reset()
    s.add_pairs_host(deep["codes"][:half], deep["quals"][:half], deep["lens"][:half])
    assert s.rmdup_finish() >= 0
    s.close()
    # without rmdup the same sample counts everything
    s = ctx.sample(idx)
    s.add_pairs_host(deep["codes"], deep["quals"], deep["lens"])
    assert np.array_equal(s.counts_host(), deep["counts_all"])
    s.close()
    idx.close()


def test_driver_rmdup_bam(deep, tmp_path):
    from tests import bamio, drvutil
    W = deep["W"]
    fa, r1, r2 = str(tmp_path / "phix.fa"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    names = drvutil.pair_names("d", deep["n"])
    drvutil.write_fastq(deep["codes"], deep["quals"], deep["lens"], names, r1, r2)
    bam, rbam, tsv, met, txt = (str(tmp_path / x) for x in ("s.bam", "s.rmdup.bam", "s.tsv", "s.metrics.txt", "s.mpileup"))
    p = drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--rmdup", 1, "--bam", bam, "--rmdup-bam", rbam, "--counts", tsv,
                            "--metrics", met, "--sample", "phix-deep", "--mpileup", txt])
    n_dup = int(deep["dup"].sum())
    assert f"rmdup: {n_dup} of {deep['n']} pairs are duplicates" in p.stderr
    full, kept = bamio.Bam(bam), bamio.Bam(rbam)
    assert len(full.records) == 2 * deep["n"]
    flagged = [r for r in full.records if r["flag"] & 0x400]
    assert len(flagged) == int((deep["marked"]["flag"] & 0x400 != 0).sum())
    assert [(r["name"], r["flag"], r["pos"]) for r in kept.records] == [(r["name"], r["flag"], r["pos"]) for r in full.records if not r["flag"] & 0x400]
    rows = [ln.rstrip("\n").split("\t") for ln in open(tsv)][1:]
    assert np.array_equal(np.array([[int(x) for x in r[4:]] for r in rows], dtype=np.int32), deep["counts"])
    assert open(met).read().splitlines()[-1].split("\t")[1:3] == [str(deep["n"]), str(n_dup)]
    refs, _ = bamio.read_bai(rbam + ".bai")
    assert len(refs) == 1
    # the text pileup is the one of the duplicate-free records (depth ~1700: lines of several kilobytes)
    from oracle import qmo_py
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    want = qmo_py.mpileup_text(ref, deep["marked"], deep["codes"], deep["quals"], deep["lens"], list(W.ref.names))
    assert open(txt, "rb").read() == want and len(want) > 10_000_000
