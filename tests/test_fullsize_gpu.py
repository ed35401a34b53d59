"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run 2 M pairs in seconds):
conservation laws that tie the count tensor to the alignment records, shard additivity (the multi-GPU decomposition),
run-to-run determinism, and the device coordinate sort on millions of records.  Every law is first checked on the CPU
oracle's own records and counts at a size the oracle handles, so the law itself is pinned before it is trusted."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
FULL = 2_000_000            # BASELINE configs[1]: 2 M pairs per sample


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


def conservation(alns, counts_rows):
    """(admitted reads, aligned M bases of admitted reads) from the records vs (sum of channel 15, sum of channel 14)"""
    f = alns["flag"].astype(np.int64)
    nc = alns["n_cigar"].astype(np.int64)
    ok = ((f & (0x4 | 0x100 | 0x200 | 0x400)) == 0) & (nc != 0) & (nc != 255) & ~(((f & 1) != 0) & ((f & 2) == 0))
    cig = alns["cigar"].astype(np.int64)
    k = np.arange(cig.shape[1])[None, :]
    is_m = ((cig & 0xf) == 0) & (k < nc[:, None])
    m_bases = ((cig >> 4) * is_m).sum(1)
    return (int(ok.sum()), int(m_bases[ok].sum())), (int(counts_rows[:, 15].sum()), int(counts_rows[:, 14].sum()))


def test_conservation_law_holds_on_the_oracle():
    from oracle import qmo_py
    from quasimodo_b200 import workloads
    W = workloads.config2(4, 5000)
    codes, quals, _, _ = W.simulate_host(0, 5000)
    lens = np.full(10000, 150, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, _, _ = qmo_py.run_sample(ref, codes, quals, lens)
    a, b = conservation(alns, counts)
    assert a == b and a[0] > 5000


@pytest.fixture(scope="module")
def full(ctx):
    """sample TA-1-1 at full size, reads simulated on the device; whole-sample run with records"""
    import torch
    from quasimodo_b200 import _lib, workloads
    W = workloads.config2(4, FULL)
    idx = ctx.index(W.ref, 31)
    dev = torch.device("cuda:0")
    g = torch.from_numpy(W.src_codes).to(dev)
    c = torch.empty((2 * FULL, 150), dtype=torch.uint8, device=dev)
    q = torch.empty_like(c)
    ctx.simulate_pairs(W, 0, FULL, g, c, q, 0)
    lens = torch.full((2 * FULL,), 150, dtype=torch.int32, device=dev)
    d_alns = torch.zeros(2 * FULL * 128, dtype=torch.uint8, device=dev)
    s = ctx.sample(idx)
    s.add_pairs(c, q, lens, d_alns=d_alns)
    counts = s.counts_host()
    pes = s.get_pestat()
    stats = s.stats()
    s.close()
    yield dict(W=W, idx=idx, c=c, q=q, lens=lens, d_alns=d_alns, counts=counts, pes=pes, stats=stats)
    idx.close()


def test_full_size_conservation(full):
    from quasimodo_b200 import _lib
    alns = full["d_alns"].cpu().numpy().view(_lib.ALN_DTYPE)
    a, b = conservation(alns, full["counts"])
    assert a == b
    assert a[0] > 3_000_000                       # most of the 4 M reads are admitted
    assert full["stats"][0] == FULL and full["stats"][1] > 5_000_000_000      # executed extension cells of the sample
    # mates point at each other
    x, y = alns[0::2], alns[1::2]
    both = ((x["flag"] & 4) == 0) & ((y["flag"] & 4) == 0)
    assert np.array_equal(x["mate_pos"][both], y["pos"][both]) and np.array_equal(y["mate_pos"][both], x["pos"][both])
    assert np.array_equal(x["tlen"][both], -y["tlen"][both])
    # every mapped record lies inside its contig
    m = (alns["flag"] & 4) == 0
    assert (alns["pos"][m] >= 0).all() and (alns["pos"][m] < np.array(full["W"].ref.lens)[alns["rid"][m]]).all()


def test_full_size_shards_add_up_and_runs_repeat(ctx, full):
    """four shards of 500 k pairs, each primed with the sample's insert-size prefix, sum to the whole-sample tensor
    (what the NCCL all-reduce adds up); a second whole-sample run reproduces the tensor bit for bit"""
    idx, c, q, lens = full["idx"], full["c"], full["q"], full["lens"]
    from quasimodo_b200 import _lib
    npre = _lib.PESTAT_PAIRS
    total = np.zeros_like(full["counts"])
    for k in range(4):
        s = ctx.sample(idx)
        lo, hi = k * FULL // 4, (k + 1) * FULL // 4
        if k:
            s.estimate_pestat(c[:2 * npre], lens[:2 * npre])
        s.add_pairs(c[2 * lo:2 * hi], q[2 * lo:2 * hi], lens[2 * lo:2 * hi], pair_id0=lo)
        assert s.get_pestat().tobytes() == full["pes"].tobytes()
        total += s.counts_host()
        s.close()
    assert np.array_equal(total, full["counts"])
    s = ctx.sample(idx)
    s.add_pairs(c, q, lens)
    assert np.array_equal(s.counts_host(), full["counts"])
    s.close()


def test_full_size_host_entry_equals_device_entry(ctx, full):
    """the host-buffer entry (chunked copies, piecewise seeding) gives the same tensor and records as the resident one"""
    import torch
    from quasimodo_b200 import _lib
    n = 600_000
    hc, hq = full["c"][:2 * n].cpu().pin_memory(), full["q"][:2 * n].cpu().pin_memory()
    hl = full["lens"][:2 * n].cpu().pin_memory()
    h_alns = np.zeros(2 * n, dtype=_lib.ALN_DTYPE)
    a = ctx.sample(full["idx"])
    a.add_pairs_host(hc, hq, hl, h_alns=h_alns)
    b = ctx.sample(full["idx"])
    d_alns = torch.zeros(2 * n * 128, dtype=torch.uint8, device="cuda")
    b.add_pairs(full["c"][:2 * n], full["q"][:2 * n], full["lens"][:2 * n], d_alns=d_alns)
    assert np.array_equal(a.counts_host(), b.counts_host())
    assert h_alns.tobytes() == d_alns.cpu().numpy().tobytes()
    a.close()
    b.close()


def test_full_size_sort_is_stable_and_ordered(ctx, full):
    import torch
    from oracle import sort_py
    from quasimodo_b200 import _lib
    n = 2 * FULL
    d_keys = torch.empty(n, dtype=torch.int64, device="cuda")
    d_perm = torch.empty(n, dtype=torch.int32, device="cuda")
    bits = C.c_int()
    L = _lib.lib()
    assert L.qm_aln_sort_keys(ctx._h, full["idx"]._h, C.c_void_p(full["d_alns"].data_ptr()), n, C.c_void_p(d_keys.data_ptr()), C.byref(bits), None) == 0
    keys_in = d_keys.clone()
    assert L.qm_sort_pairs(ctx._h, C.c_void_p(d_keys.data_ptr()), C.c_void_p(d_perm.data_ptr()), n, bits.value, None) == 0
    torch.cuda.synchronize()
    perm = d_perm.to(torch.int64)
    assert bool((d_keys[1:] >= d_keys[:-1]).all())                                   # ordered
    assert bool((keys_in[perm] == d_keys).all())                                      # a gather of the input
    assert int(torch.bincount(perm, minlength=n).max()) == 1                          # a permutation
    tie = d_keys[1:] == d_keys[:-1]
    assert bool((perm[1:][tie] > perm[:-1][tie]).all())                               # ties in input order
    # and the order is samtools' order of the records themselves
    alns = full["d_alns"].cpu().numpy().view(_lib.ALN_DTYPE)
    sk = sort_py.samtools_keys(alns)[d_perm.cpu().numpy().view(np.uint32)]
    assert (sk[1:] >= sk[:-1]).all()


def test_call_larger_than_one_internal_chunk(ctx, full):
    """2.3 M pairs in ONE call (the library walks it in chunks of 2^21 pairs, seeding in batches of 4 M reads) equals the sum of
    two calls split at an odd boundary, records included"""
    import torch
    from quasimodo_b200 import _lib
    n = 2_300_000
    W, idx = full["W"], full["idx"]
    dev = torch.device("cuda:0")
    g = torch.from_numpy(W.src_codes).to(dev)
    c = torch.empty((2 * n, 150), dtype=torch.uint8, device=dev)
    q = torch.empty_like(c)
    ctx.simulate_pairs(W, 0, n, g, c, q, 0)
    lens = torch.full((2 * n,), 150, dtype=torch.int32, device=dev)
    whole_alns = torch.zeros(2 * n * 128, dtype=torch.uint8, device=dev)
    s = ctx.sample(idx)
    s.add_pairs(c, q, lens, d_alns=whole_alns)
    whole = s.counts_host()
    assert s.stats()[0] == n
    s.close()
    cut = 1_000_003
    npre = _lib.PESTAT_PAIRS
    total = np.zeros_like(whole)
    parts = []
    for lo, hi in ((0, cut), (cut, n)):
        s = ctx.sample(idx)
        if lo:
            s.estimate_pestat(c[:2 * npre], lens[:2 * npre])
        d_alns = torch.zeros(2 * (hi - lo) * 128, dtype=torch.uint8, device=dev)
        s.add_pairs(c[2 * lo:2 * hi], q[2 * lo:2 * hi], lens[2 * lo:2 * hi], pair_id0=lo, d_alns=d_alns)
        total += s.counts_host()
        parts.append(d_alns)
        s.close()
    assert np.array_equal(total, whole)
    assert torch.equal(torch.cat(parts), whole_alns)


def _run_whole(ctx, W, n_pairs, L, opt, calls=1, records=True):
    """simulate n_pairs on the device and run them through one qm_sample in `calls` library calls -> (counts, alns or None, stats, pes, tensors)"""
    import torch
    from quasimodo_b200 import _lib
    dev = torch.device("cuda:0")
    g = torch.from_numpy(W.src_codes).to(dev)
    c = torch.empty((2 * n_pairs, L), dtype=torch.uint8, device=dev)
    q = torch.empty_like(c)
    step = 1 << 21
    for o in range(0, n_pairs, step):
        m = min(step, n_pairs - o)
        ctx.simulate_pairs(W, o, m, g, c[2 * o:2 * (o + m)], q[2 * o:2 * (o + m)], 0)
    lens = torch.full((2 * n_pairs,), L, dtype=torch.int32, device=dev)
    idx = ctx.index(W.ref, 31)
    s = ctx.sample(idx, opt)
    d_alns = torch.zeros(2 * n_pairs * 128, dtype=torch.uint8, device=dev) if records else None
    per = (n_pairs + calls - 1) // calls
    for k in range(calls):
        lo, hi = k * per, min(n_pairs, (k + 1) * per)
        s.add_pairs(c[2 * lo:2 * hi], q[2 * lo:2 * hi], lens[2 * lo:2 * hi], pair_id0=lo,
                    d_alns=None if d_alns is None else d_alns[2 * lo * 128:2 * hi * 128])
    counts, stats, pes = s.counts_host(), s.stats(), s.get_pestat()
    s.close()
    alns = None if d_alns is None else d_alns.cpu().numpy().view(_lib.ALN_DTYPE)
    return counts, alns, stats, pes, (idx, c, q, lens)


def test_full_size_config5_long_reads_with_indels(ctx):
    """BASELINE configs[4] at size: 2 M pairs of 2 x 250 bp with indels, band 200.  Conservation between records and tensor,
    deletion / insertion events between CIGARs and channels 12 / 13 / 5 / 11, two shards add up to the whole."""
    from quasimodo_b200 import _lib, workloads
    n = 2_000_000
    W = workloads.config5(n)
    opt = _lib.default_opt()
    opt.w = 200
    counts, alns, stats, pes, (idx, c, q, lens) = _run_whole(ctx, W, n, 250, opt)
    a, b = conservation(alns, counts)
    assert a == b and a[0] > 3_000_000
    assert stats[0] == n and stats[1] > 40_000_000_000             # ~29 k extension cells per pair
    f = alns["flag"].astype(np.int64)
    nc = alns["n_cigar"].astype(np.int64)
    ok = ((f & (0x4 | 0x100 | 0x200 | 0x400)) == 0) & (nc != 0) & (nc != 255) & ~(((f & 1) != 0) & ((f & 2) == 0))
    cig = alns["cigar"].astype(np.int64)[ok]
    live = np.arange(cig.shape[1])[None, :] < nc[ok][:, None]
    after_m = np.cumsum(((cig & 0xf) == 0) & live, axis=1) > 0           # the event channels need an aligned base in front
    n_ins, n_del = int((((cig & 0xf) == 1) & live & after_m).sum()), int((((cig & 0xf) == 2) & live & after_m).sum())
    del_bases = int(((cig >> 4) * (((cig & 0xf) == 2) & live)).sum())
    assert n_ins > 50_000 and n_del > 50_000
    assert int(counts[:, 12].sum()) == n_ins and int(counts[:, 13].sum()) == n_del
    assert int(counts[:, 5].sum() + counts[:, 11].sum()) == del_bases
    # two shards, the second primed with the sample's insert-size prefix
    total = np.zeros_like(counts)
    npre = _lib.PESTAT_PAIRS
    for k in range(2):
        s = ctx.sample(idx, opt)
        lo, hi = k * n // 2, (k + 1) * n // 2
        if k:
            s.estimate_pestat(c[:2 * npre], lens[:2 * npre])
        s.add_pairs(c[2 * lo:2 * hi], q[2 * lo:2 * hi], lens[2 * lo:2 * hi], pair_id0=lo)
        assert s.get_pestat().tobytes() == pes.tobytes()
        total += s.counts_host()
        s.close()
    assert np.array_equal(total, counts)
    idx.close()


def test_config4_deep_sample_in_several_calls(ctx):
    """BASELINE configs[3]'s sample (TM-1-50, one reference, 60,000x when whole) at 6 M pairs: three library calls of 2 M pairs give
    the tensor of one call of 6 M (the batching of a 50 M-pair sample does not show in the result), conservation holds, and the depth
    is what the pair count implies"""
    from quasimodo_b200 import _lib, workloads
    n = 6_000_000
    W = workloads.config4(50_000_000)
    opt = _lib.default_opt()
    c3, alns, stats3, pes3, (idx, c, q, lens) = _run_whole(ctx, W, n, 150, opt, calls=3)
    a, b = conservation(alns, c3)
    assert a == b and a[0] > 10_000_000
    del alns
    s = ctx.sample(idx, opt)
    s.add_pairs(c, q, lens)
    assert np.array_equal(s.counts_host(), c3)
    assert s.stats() == stats3 and s.get_pestat().tobytes() == pes3.tobytes()
    s.close()
    depth = c3[:, 14].sum() / W.ref.total
    assert 0.8 * 2 * n * 150 / W.ref.total < depth <= 2 * n * 150 / W.ref.total
    idx.close()


def test_full_size_config3_contaminated_sample(ctx):
    """BASELINE configs[2] at size: 1 M pairs of AD169:Merlin 1:10 + 5 % PhiX + 5 % E. coli against the 4.89 Mb three-contig index.
    Conservation between records and tensor; the reads of each source land on its contig (the decontamination rule's premise);
    two library calls give the tensor of one."""
    from quasimodo_b200 import _lib, workloads
    n = 1_000_000
    W = workloads.config3(n)
    opt = _lib.default_opt()
    counts, alns, stats, pes, (idx, c, q, lens) = _run_whole(ctx, W, n, 150, opt)
    a, b = conservation(alns, counts)
    assert a == b and a[0] > 1_800_000
    m = (alns["flag"] & 4) == 0
    share = np.bincount(alns["rid"][m], minlength=3) / m.sum()
    assert 0.86 < share[0] < 0.93 and 0.03 < share[1] < 0.07 and 0.03 < share[2] < 0.07, share
    assert (alns["pos"][m] < np.array(W.ref.lens)[alns["rid"][m]]).all()
    # depth per contig follows the read shares: channel 14 summed inside each contig
    off = np.concatenate([[0], np.cumsum(W.ref.lens)])
    per = np.array([counts[off[k]:off[k + 1], 14].sum() for k in range(3)], dtype=np.float64)
    assert abs(per[0] / per.sum() - share[0]) < 0.03
    c2, _, stats2, pes2, t2 = _run_whole(ctx, W, n, 150, opt, calls=2, records=False)
    assert np.array_equal(c2, counts) and stats2 == stats and pes2.tobytes() == pes.tobytes()
    t2[0].close()
    idx.close()
