"""Text pileup (samtools mpileup format) produced on the device against the oracle's
restatement, byte for byte, on the BASELINE configs (indels: cfg5; several contigs: cfg3), plus format properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


def run(ctx, W, n, popt=None):
    import torch
    from quasimodo_b200 import _lib
    from oracle import qmo_py
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    opt_o = qmo_py.default_opt(); opt_o.w = W.w
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, _, _, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=opt_o)
    names = [f"contig{i}" for i in range(len(W.ref.lens))]
    po_o = po_g = None
    if popt:
        import ctypes
        po_o, po_g = qmo_py.PileupOpt(), _lib.default_pileup_opt()
        qmo_py.lib().qmo_pileup_opt_default(ctypes.byref(po_o))
        for k, v in popt.items():
            setattr(po_o, k, v)
            setattr(po_g, k, v)
    o_text = qmo_py.mpileup_text(ref, alns, codes, quals, lens, names, popt=po_o)
    dev = torch.device("cuda:0")
    idx = ctx.index(W.ref, 31)
    g_text = ctx.mpileup_text(idx, torch.from_numpy(alns.view(np.uint8).reshape(-1)).to(dev), torch.from_numpy(codes).to(dev),
                              torch.from_numpy(quals).to(dev), torch.from_numpy(lens).to(dev), names, popt=po_g)
    idx.close()
    return g_text, o_text, alns


@pytest.mark.parametrize("name,n", [("cfg1", 3000), ("cfg5", 2000), ("cfg3", 3000), ("cfg2", 20000)])
def test_text_pileup_equals_oracle(ctx, name, n):
    from tests.test_pipeline_gpu import _workload
    g, o, _ = run(ctx, _workload(name, n), n)
    assert len(o) > 100000
    if g != o:
        gl, ol = g.split(b"\n"), o.split(b"\n")
        assert len(gl) == len(ol), (len(gl), len(ol))
        bad = [i for i in range(len(ol)) if gl[i] != ol[i]]
        assert not bad, (len(bad), gl[bad[0]][:200], ol[bad[0]][:200])
    assert g == o


@pytest.mark.parametrize("popt", [dict(min_bq=25, min_mapq=20), dict(count_orphans=1, ignore_overlaps=1), dict(min_bq=0)])
def test_text_pileup_options(ctx, popt):
    """-Q / -q / -A / -x off their defaults"""
    from tests.test_pipeline_gpu import _workload
    g, o, _ = run(ctx, _workload("cfg5", 1500), 1500, popt=popt)
    assert len(o) > 100000 and g == o
    if popt.get("min_bq") == 25:
        # columns whose every base fails -Q: depth 0 and "*" for both strings (samtools 1.9 bam_plcmd.c), never empty fields
        assert b"\t0\t*\t*\n" in g and b"\t\t" not in g


def test_text_pileup_format(ctx):
    from tests.test_pipeline_gpu import _workload
    g, _, alns = run(ctx, _workload("cfg5", 1500), 1500)
    lines = [l.split(b"\t") for l in g.split(b"\n") if l]
    assert all(len(f) == 6 for f in lines)
    assert all(int(f[3]) == len(f[5]) or (f[3] == b"0" and f[4] == b"*" and f[5] == b"*") for f in lines)
    pos = [int(f[1]) for f in lines]
    assert pos == sorted(pos)
    text = b"".join(f[4] for f in lines)
    # every admitted read opens and closes once, unless the quality at its first / last column fails the filter
    admitted = ((alns["flag"] & 4) == 0) & ((alns["flag"] & 2) != 0)
    assert 0.9 * admitted.sum() < text.count(b"^") <= admitted.sum()
    assert b"+" in text and b"-" in text and b"*" in text


def test_text_pileup_empty(ctx):
    """no admitted read -> no line"""
    import torch
    from quasimodo_b200 import workloads, _lib
    W = workloads.config1(64)
    codes, quals, _, _ = W.simulate_host(0, 64)
    lens = np.full(128, 150, np.int32)
    alns = np.zeros(128, dtype=_lib.ALN_DTYPE)
    alns["flag"] = 4; alns["rid"] = -1; alns["n_cigar"] = 0
    dev = torch.device("cuda:0")
    idx = ctx.index(W.ref, 31)
    g = ctx.mpileup_text(idx, torch.from_numpy(alns.view(np.uint8).reshape(-1)).to(dev), torch.from_numpy(codes).to(dev),
                         torch.from_numpy(quals).to(dev), torch.from_numpy(lens).to(dev), ["x"])
    idx.close()
    assert g == b""
