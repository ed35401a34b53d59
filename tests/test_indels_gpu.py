"""GPU parity of the indel allele table: every insertion / deletion of an admitted read's CIGAR tallied per
(anchor, type, length, inserted bases) in a device hash table, against the CPU restatement (oracle/qmo_pileup.c qmo_indels);
and the merge step of the multi-GPU gather: two half-sample tables merged equal the whole sample's table."""
import ctypes as C

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
def case():
    from oracle import qmo_py
    from quasimodo_b200 import workloads
    n = 6000
    W = workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 250, np.int32)
    opt = qmo_py.default_opt()
    opt.w = 200
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, _, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)
    return dict(W=W, n=n, codes=codes, quals=quals, lens=lens, alns=alns, counts=counts, ref=ref,
                want=qmo_py.indels(ref, alns, codes, lens))


def as_rows(rec):
    key = rec["key"].astype(np.uint64)
    return np.stack([rec["rid"], rec["pos"], rec["len"], rec["type"].astype(np.int32), rec["has_n"].astype(np.int32),
                     rec["seq"].astype(np.int64).astype(np.int32), rec["n_fwd"], rec["n_rev"],
                     (key & np.uint64(0xffffffff)).astype(np.uint32).view(np.int32), (key >> np.uint64(32)).astype(np.uint32).view(np.int32)], 1)


def test_indel_table_matches_oracle(ctx, case):
    from quasimodo_b200 import _lib
    opt = _lib.default_opt()
    opt.w = 200
    idx = ctx.index(case["W"].ref, 31)
    s = ctx.sample(idx, opt)
    s.add_pairs_host(case["codes"], case["quals"], case["lens"])
    got = s.indels()
    want = case["want"]
    assert len(want) > 300 and (want[:, 3] == 0).any() and (want[:, 3] == 1).any()
    assert np.array_equal(as_rows(got), want)
    # the table and the dense event channels tell the same story: events per anchor
    cnt = s.counts_host()
    assert int(got["n_fwd"].sum() + got["n_rev"].sum()) == int(cnt[:, 12].sum() + cnt[:, 13].sum())
    s.reset()
    assert len(s.indels()) == 0
    s.close()
    idx.close()


def test_half_sample_tables_merge_to_the_whole(ctx, case):
    """what the multi-GPU gather does after the all-gather: another rank's records are added into the local table"""
    import torch
    from quasimodo_b200 import _lib
    opt = _lib.default_opt()
    opt.w = 200
    idx = ctx.index(case["W"].ref, 31)
    h = case["n"] // 2 * 2
    a, b = ctx.sample(idx, opt), ctx.sample(idx, opt)
    a.add_pairs_host(case["codes"][:h], case["quals"][:h], case["lens"][:h])
    b.set_pestat(a.get_pestat())                     # one insert-size model per sample
    b.add_pairs_host(case["codes"][h:], case["quals"][h:], case["lens"][h:], pair_id0=h // 2)
    rb = b.indels()
    assert 0 < len(rb) < len(case["want"])
    d = torch.from_numpy(rb.view(np.uint8).reshape(-1).copy()).cuda()
    L = _lib.lib()
    assert L.qm_indel_table_merge(L.qm_sample_indel_table(a._h), C.c_void_p(d.data_ptr()), len(rb), None) == 0
    assert np.array_equal(as_rows(a.indels()), case["want"])
    a.close(); b.close(); idx.close()
