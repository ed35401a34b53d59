"""GPU tests of the FM-index seeder: the index rebuilt by the library equals bwa's own files shipped in the
reference (digests, tests/golden/fm_digests.json); with QM_F_FM_SEEDS the seeds are the oracle's restatement of bwa's seeding,
seed for seed in bwa's order; and the whole path -- regions, records, counts -- agrees with the oracle run on the same seeds."""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fm_digests.json")))


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("stem", ["Phix", "Merlin", "AD169"])
def test_library_builds_bwas_index_files(ctx, stem):
    from quasimodo_b200 import genomes
    G = genomes.load(stem)
    idx = ctx.index(G, 31)
    idx.build_fm()
    b, s = idx.fm_export()
    assert hashlib.sha256(b).hexdigest() == GOLD[stem]["bwt"]["sha256"] and len(b) == GOLD[stem]["bwt"]["bytes"]
    assert hashlib.sha256(s).hexdigest() == GOLD[stem]["sa"]["sha256"] and len(s) == GOLD[stem]["sa"]["bytes"]
    idx.attach_bwa(b, s)                                  # and the file route takes them back
    assert idx.fm_export() == (b, s)
    from quasimodo_b200 import QmError
    other = genomes.load("Phix" if stem != "Phix" else "Merlin")
    odx = ctx.index(other, 31)
    with pytest.raises(QmError):
        odx.attach_bwa(b, s)                              # another genome's index is refused
    odx.close()
    idx.close()


@pytest.mark.parametrize("cfg,n,flags", [("cfg2", 3000, 2), ("cfg5", 1500, 2), ("cfg2", 1500, 6)])
def test_fm_seeds_and_pipeline_match_oracle(ctx, cfg, n, flags):
    import torch
    from oracle import qmo_py
    from quasimodo_b200 import _lib, workloads
    W = workloads.config2(3, n) if cfg == "cfg2" else workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    L = W.params.read_len
    lens = np.full(2 * n, L, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    ref.set_fm(qmo_py.FmIndex(codes=W.ref.codes), max_mem_intv=0 if flags & 4 else 20)
    oo = qmo_py.default_opt()
    oo.w = W.w
    oo.flags |= qmo_py.F_FM_SEEDS
    o = qmo_py.align_se(ref, codes, lens, opt=oo)
    alns, counts, cells, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=oo)
    og = _lib.default_opt()
    og.w = W.w
    og.flags |= flags
    idx = ctx.index(W.ref, 31)
    from quasimodo_b200 import QmError
    dev = torch.device("cuda:0")
    d_codes, d_lens = torch.from_numpy(codes).to(dev), torch.from_numpy(lens).to(dev)
    with pytest.raises(QmError):
        ctx.collect_seeds(idx, d_codes, d_lens, opt=og)   # no FM-index attached yet
    idx.build_fm()
    seeds, n_seeds = ctx.collect_seeds(idx, d_codes, d_lens, opt=og)
    torch.cuda.synchronize()
    g_n = n_seeds.cpu().numpy()
    g_s = seeds.cpu().numpy().view(_lib.SEED_DTYPE).reshape(2 * n, _lib.MAX_SEEDS)
    assert np.array_equal(g_n, o["n_seeds"])
    assert g_n.mean() > (1.2 if flags & 4 else 2.0)
    for r in range(2 * n):
        k = g_n[r]
        for f in ("rbeg", "qbeg", "len"):
            assert np.array_equal(g_s[r, :k][f], o["seeds"][r, :k][f]), (r, f)
    s = ctx.sample(idx, og)
    h_alns = np.zeros(2 * n, dtype=_lib.ALN_DTYPE)
    s.add_pairs_host(codes, quals, lens, h_alns=h_alns)
    assert s.stats() == (n, cells)
    for f in ("rid", "pos", "flag", "mapq", "n_cigar", "score", "sub", "nm", "mate_rid", "mate_pos", "tlen"):
        assert np.array_equal(h_alns[f], alns[f]), f
    assert np.array_equal(s.counts_host(), counts)
    s.close()
    idx.close()


def test_driver_with_bwas_index_files(ctx, tmp_path):
    """qm_driver sample --bwa-index PREFIX: the index files as `bwa index` writes them (here exported by the library, byte-equal
    to the reference's own) drive the seeding; records equal the oracle's FM-seeded run"""
    from oracle import qmo_py, sort_py
    from quasimodo_b200 import genomes, workloads
    from tests import bamio, drvutil
    n = 2000
    W = workloads.config1(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    idx = ctx.index(W.ref, 31)
    idx.build_fm()
    b, s = idx.fm_export()
    idx.close()
    fa, r1, r2 = str(tmp_path / "Merlin.fa"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    open(fa + ".bwt", "wb").write(b)
    open(fa + ".sa", "wb").write(s)
    names = drvutil.pair_names("f", n)
    drvutil.write_fastq(codes, quals, lens, names, r1, r2)
    bam = str(tmp_path / "s.bam")
    drvutil.run_driver(["sample", "--ref", fa, "--bwa-index", fa, "--r1", r1, "--r2", r2, "--bam", bam])
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    ref.set_fm(qmo_py.FmIndex(bwt=b, sa=s))
    oo = qmo_py.default_opt()
    oo.flags |= qmo_py.F_FM_SEEDS
    alns, _, _, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=oo)
    case = dict(bam=bamio.Bam(bam), alns=alns, perm=sort_py.sort_perm(alns), names=names, codes=codes, quals=quals, lens=lens, W=W)
    assert drvutil.check_bam_records(case) > n
