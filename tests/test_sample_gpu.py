"""GPU parity of the rule-level entry (qm_sample_*): device batches, host batches, sharded batches and the
SNP-call / TP-FP stages, against the CPU oracle and the golden files of the reference's own script."""
import filecmp
import os
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval")


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def case(ctx):
    """cfg2 sample TA-1-10, 6000 pairs: oracle results + device index"""
    from quasimodo_b200 import workloads
    from oracle import qmo_py
    n = 6000
    W = workloads.config2(6, n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, cells, pes = qmo_py.run_sample(ref, codes, quals, lens)
    idx = ctx.index(W.ref, 31)
    yield dict(W=W, n=n, codes=codes, quals=quals, lens=lens, ref=ref, alns=alns, counts=counts, cells=cells, pes=pes, idx=idx)
    idx.close()


def test_sample_device_batch(ctx, case):
    import torch
    from quasimodo_b200 import _lib
    s = ctx.sample(case["idx"])
    dev = torch.device("cuda:0")
    d_codes, d_quals = torch.from_numpy(case["codes"]).to(dev), torch.from_numpy(case["quals"]).to(dev)
    d_lens = torch.from_numpy(case["lens"]).to(dev)
    d_alns = torch.zeros(2 * case["n"] * 128, dtype=torch.uint8, device=dev)
    s.add_pairs(d_codes, d_quals, d_lens, d_alns=d_alns)
    assert np.array_equal(s.counts_host(), case["counts"])
    assert s.stats() == (case["n"], case["cells"])
    assert s.get_pestat().tobytes() == case["pes"].tobytes()
    g = d_alns.cpu().numpy().view(_lib.ALN_DTYPE)
    for f in ("rid", "pos", "flag", "mapq", "n_cigar", "nm", "tlen"):
        assert np.array_equal(g[f], case["alns"][f]), f
    # reset + rerun gives the same tensor; two runs without reset give twice the counts
    s.add_pairs(d_codes, d_quals, d_lens)
    assert np.array_equal(s.counts_host(), 2 * case["counts"])
    s.reset()
    s.add_pairs(d_codes, d_quals, d_lens)
    assert np.array_equal(s.counts_host(), case["counts"])
    s.close()


def test_sample_host_batch(ctx, case):
    from quasimodo_b200 import _lib
    s = ctx.sample(case["idx"])
    h_alns = np.zeros(2 * case["n"], dtype=_lib.ALN_DTYPE)
    s.add_pairs_host(case["codes"], case["quals"], case["lens"], h_alns=h_alns)
    assert np.array_equal(s.counts_host(), case["counts"])
    for f in ("rid", "pos", "flag", "mapq", "n_cigar", "nm", "tlen"):
        assert np.array_equal(h_alns[f], case["alns"][f]), f
    s.close()


def test_sample_sharded_equals_whole(ctx, case):
    """two shards (the second one primed with the sample's insert-size prefix) sum to the whole sample's counts --
    the multi-GPU decomposition (counts are added by the all-reduce)"""
    import torch
    dev = torch.device("cuda:0")
    n, half = case["n"], case["n"] // 2
    d_codes, d_quals = torch.from_numpy(case["codes"]).to(dev), torch.from_numpy(case["quals"]).to(dev)
    d_lens = torch.from_numpy(case["lens"]).to(dev)
    a, b = ctx.sample(case["idx"]), ctx.sample(case["idx"])
    # rank 0: the whole prefix is its own first pairs only when the shard holds >= PESTAT_PAIRS pairs; here the
    # sample is smaller than the prefix, so both shards estimate from the whole sample's pairs
    a.estimate_pestat(d_codes, d_lens)
    b.estimate_pestat(d_codes, d_lens)
    a.add_pairs(d_codes[:2 * half], d_quals[:2 * half], d_lens[:2 * half], pair_id0=0)
    b.add_pairs(d_codes[2 * half:], d_quals[2 * half:], d_lens[2 * half:], pair_id0=half)
    assert np.array_equal(a.counts_host() + b.counts_host(), case["counts"])
    assert a.get_pestat().tobytes() == case["pes"].tobytes()
    a.close()
    b.close()


def test_call_snps_thresholds(ctx, case):
    """the caller's records follow its stated rule on the oracle's counts; sorted; REF is the reference base"""
    from quasimodo_b200 import _lib
    s = ctx.sample(case["idx"])
    s.add_pairs_host(case["codes"], case["quals"], case["lens"])
    calls = s.call_snps()
    cnt = case["counts"].astype(np.int64)
    ad = cnt[:, 0:4] + cnt[:, 6:10]
    tot = ad.sum(1)
    refb = case["W"].ref.codes
    want = []
    for p in np.nonzero(cnt[:, 14] >= 10)[0]:
        for b in range(4):
            if b != refb[p] and ad[p, b] >= 2 and ad[p, b] >= np.float32(0.01).astype(np.float64) * tot[p]:
                want.append((p, b))
    got = [(int(c["pos"]), int(c["alt"])) for c in calls]
    assert got == want and len(got) > 50
    assert all(int(c["ref"]) == refb[int(c["pos"])] for c in calls)
    assert all(int(c["ad_alt_f"]) + int(c["ad_alt_r"]) == ad[int(c["pos"]), int(c["alt"])] for c in calls)
    s.close()


@pytest.mark.parametrize("sample", ["TM-1-1", "TA-1-0"])
def test_extract_tp_fp_matches_reference_script(ctx, tmp_path, sample):
    """product evaluate.extract_tp_fp_snp (CUDA matcher) == the reference script's own output files"""
    from quasimodo_b200 import evaluate
    src = os.path.join(GOLD, sample)
    d = tmp_path / sample
    os.makedirs(d)
    name = f"{sample}.Merlin.bcftools"
    shutil.copy(os.path.join(src, name + ".vcf"), d)
    evaluate.extract_tp_fp_snp(ctx, str(d / (name + ".vcf")), os.path.join(src, "truth.vcf"))
    assert filecmp.cmp(d / (name + ".filtered.vcf"), os.path.join(src, name + ".filtered.vcf"), shallow=False)
    assert filecmp.cmp(d / "fp" / (name + ".fp.vcf"), os.path.join(src, "fp", name + ".fp.vcf"), shallow=False)
    if sample == "TM-1-1":
        assert filecmp.cmp(d / "tp" / (name + ".tp.vcf"), os.path.join(src, "tp", name + ".tp.vcf"), shallow=False)


def test_custom_dataset_evaluator_matches_reference_script(ctx, tmp_path):
    """product evaluate.extract_tp_fp_custom_snp (truth = show-snps rows, CUDA matcher) == the files the reference's own
    script writes in its "custom" mode; the benchmark-table row == the restatement of custom_snp_benchmark.R"""
    from quasimodo_b200 import evaluate
    from oracle import eval_py
    src = os.path.join(GOLD, "custom")
    d = tmp_path / "custom"
    os.makedirs(d)
    n, tp, fp = evaluate.extract_tp_fp_custom_snp(ctx, os.path.join(src, "mysample.calls.vcf"), os.path.join(src, "genome_diff.snps"), str(d), "mycaller")
    assert tp > 50 and fp > 50 and n == tp + fp
    for rel in ("mycaller.filtered.vcf", os.path.join("fp", "mycaller.fp.vcf"), os.path.join("tp", "mycaller.tp.vcf")):
        assert filecmp.cmp(d / rel, os.path.join(src, rel), shallow=False), rel
    got = evaluate.custom_performance_row(ctx, str(d / "mycaller.filtered.vcf"), os.path.join(src, "genome_diff.snps"), "mycaller")
    assert got == eval_py.custom_performance_row(os.path.join(src, "mycaller.filtered.vcf"), os.path.join(src, "genome_diff.snps"), "mycaller")


def test_performance_row_matches_oracle(ctx):
    from quasimodo_b200 import evaluate
    from oracle import eval_py
    src = os.path.join(GOLD, "TM-1-1")
    f = os.path.join(src, "TM-1-1.Merlin.bcftools.filtered.vcf")
    truth = {"TM": eval_py.make_snp_vector(os.path.join(src, "truth.vcf"))}
    want = eval_py.performance_row(f, truth, ["TM-1-1"])
    got, fn = evaluate.performance_row(ctx, f, {"TM": os.path.join(src, "truth.vcf")}, ["TM-1-1"])
    assert got == want
    assert fn == len(set(truth["TM"]) - set(eval_py.make_snp_vector(f)))


def test_eval_match_large_random(ctx):
    """matcher on 200k random keys with duplicates vs numpy set membership"""
    from quasimodo_b200 import evaluate
    rng = np.random.default_rng(3)
    t = rng.integers(1, 1 << 30, 150_000).astype(np.uint64)
    c = np.concatenate([rng.choice(t, 60_000), rng.integers(1, 1 << 30, 140_000).astype(np.uint64)])
    cf, tf = evaluate.match_keys(ctx, c, t)
    assert np.array_equal(cf.astype(bool), np.isin(c, t))
    assert np.array_equal(tf.astype(bool), np.isin(t, c))
    cf, tf = evaluate.match_keys(ctx, c[:0], t)
    assert len(cf) == 0 and not tf.any()


def test_eval_calls_totals_match_the_key_matcher(ctx):
    """qm_eval_calls (keys from the call records + both membership passes + TP/FP/FN totals, all on the device) against
    qm_eval_match_host on keys packed on the host"""
    import ctypes as C
    import torch
    from quasimodo_b200 import _lib, evaluate
    rng = np.random.default_rng(11)
    n, nt = 30_000, 25_000
    calls = np.zeros(n, dtype=_lib.CALL_DTYPE)
    calls["pos"] = np.sort(rng.integers(0, 200_000, n))
    calls["ref"] = rng.integers(0, 4, n)
    calls["alt"] = (calls["ref"] + rng.integers(1, 4, n)) & 3
    keys = ((calls["pos"].astype(np.uint64) + 1) << 8) | (calls["ref"].astype(np.uint64) << 4) | calls["alt"].astype(np.uint64)
    truth = np.concatenate([rng.choice(keys, 9_000, replace=False), rng.integers(1 << 8, 1 << 26, nt - 9_000).astype(np.uint64)])
    cf, tf = evaluate.match_keys(ctx, keys, truth)
    d_calls = torch.from_numpy(calls.view(np.uint8)).cuda()
    d_truth = torch.from_numpy(truth.view(np.int64)).cuda()
    d_flags = torch.zeros(n, dtype=torch.uint8, device="cuda")
    out = (C.c_int64 * 3)()
    for flags in (d_flags, None):
        rc = _lib.lib().qm_eval_calls(ctx._h, C.c_void_p(d_calls.data_ptr()), n, C.c_void_p(d_truth.data_ptr()), nt,
                                      C.c_void_p(flags.data_ptr()) if flags is not None else None, out, None)
        assert rc == 0
        assert list(out) == [int(cf.sum()), n - int(cf.sum()), nt - int(tf.sum())]
    assert np.array_equal(d_flags.cpu().numpy(), cf)
    rc = _lib.lib().qm_eval_calls(ctx._h, None, 0, C.c_void_p(d_truth.data_ptr()), nt, None, out, None)
    assert rc == 0 and list(out) == [0, 0, nt]


def test_packed_host_entry_equals_the_byte_entry(ctx):
    """qm_sample_add_pairs_host_packed (2-bit bases + N mask over the link, expanded on the device) against the 1-byte-per-base
    entry: same records, same counts -- ragged read lengths, N bases, a stride that is not a multiple of 8"""
    from quasimodo_b200 import _lib, workloads
    from quasimodo_b200.api import pack_reads
    n = 5000
    W = workloads.config1(n)
    codes, quals, _, _ = W.simulate_host(0, n, stride=157)
    lens = np.full(2 * n, 150, np.int32)
    rng = np.random.default_rng(4)
    for r in rng.integers(0, 2 * n, 200):
        lens[r] = int(rng.integers(40, 150))
        codes[r, lens[r]:] = 4
    idx = ctx.index(W.ref, 31)
    a, b = ctx.sample(idx), ctx.sample(idx)
    ha, hb = np.zeros(2 * n, dtype=_lib.ALN_DTYPE), np.zeros(2 * n, dtype=_lib.ALN_DTYPE)
    a.add_pairs_host(codes, quals, lens, h_alns=ha)
    b2, nm = pack_reads(codes)
    assert b2.shape == (2 * n, 40) and nm.shape == (2 * n, 20) and nm.any()
    b.add_pairs_host_packed(b2, nm, quals, lens, h_alns=hb)
    assert ha.tobytes() == hb.tobytes()
    assert np.array_equal(a.counts_host(), b.counts_host()) and a.stats() == b.stats()
    a.close(); b.close(); idx.close()
