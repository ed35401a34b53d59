"""GPU parity (bit-exact) of the whole read-level path, stage by stage, through the C-ABI against the CPU
oracle on the same simulated read pairs: seeds -> regions (+ executed extension cells) -> insert-size
model -> alignment records (pos, flag, MAPQ, CIGAR, NM, mate fields) -> pileup count tensor.
Integer / index work => every field must be identical; the insert-size moments are doubles computed by
the same host code path (libm) and are compared exactly as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


def _workload(name, n):
    from quasimodo_b200 import workloads
    if name == "cfg1":
        return workloads.config1(n)
    if name == "cfg2":
        return workloads.config2(6, n)
    if name == "cfg2-rescue-heavy":           # TA-50-1: TB40E reads on AD169, 5.6 % of the pairs go through mate rescue
        return workloads.config2(1, n)
    if name == "cfg3":
        return workloads.config3(n)
    if name == "cfg4":
        return workloads.config4(n)
    if name == "cfg5":
        return workloads.config5(n)
    raise KeyError(name)


def run_both(ctx, W, n_pairs, pair0=0, flags=0, scoring=None, all_orientations=False, k=31, ragged=False, popt=None):
    import torch
    from quasimodo_b200 import _lib
    from oracle import qmo_py
    codes, quals, _, _ = W.simulate_host(pair0, n_pairs)
    lens = np.full(2 * n_pairs, W.params.read_len, np.int32)
    if ragged:                      # reads trimmed to 35 .. full length (the tail of the row is padding); a few degenerate ones
        lens = np.random.default_rng(11).integers(35, W.params.read_len + 1, 2 * n_pairs).astype(np.int32)
        lens[[5, 18, 40, 77, 100, 101]] = [0, 1, 20, 30, 31, 32]
        for r in range(2 * n_pairs):
            codes[r, lens[r]:] = 4
    opt_o = qmo_py.default_opt()
    opt_o.w = W.w
    opt_o.flags = flags
    opt_g = _lib.default_opt()
    opt_g.w = W.w
    opt_g.flags = flags
    for name, v in (scoring or {}).items():
        setattr(opt_o, name, v)
        setattr(opt_g, name, v)
    # ---- oracle ----
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=k)
    o = qmo_py.align_se(ref, codes, lens, opt=opt_o)
    o_regs_se, o_nr_se = o["regs"].copy(), o["n_regs"].copy()
    o_pes = qmo_py.pestat(ref, o["regs"], o["n_regs"], opt=opt_o)
    o_regs_resc, o_nr_resc = o["regs"].copy(), o["n_regs"].copy()
    resc_pes = o_pes.copy()
    if all_orientations:            # every orientation gets the FR model: all four window shapes of mem_matesw are searched
        for d in (0, 2, 3):
            resc_pes[d] = o_pes[1]
    o_resc_stats = qmo_py.mate_rescue(ref, codes, lens, o_regs_resc, o_nr_resc, resc_pes, opt=opt_o)
    o_alns = qmo_py.pair_and_finish(ref, codes, lens, o["regs"], o["n_regs"], o_pes, pair_id0=pair0, opt=opt_o)
    po_o = None
    if popt:
        po_o = qmo_py.PileupOpt()
        qmo_py.lib().qmo_pileup_opt_default(__import__("ctypes").byref(po_o))
        for kk, v in popt.items():
            setattr(po_o, kk, v)
    o_counts = qmo_py.pileup(ref, o_alns, codes, quals, lens, popt=po_o)
    # ---- device ----
    dev = torch.device("cuda:0")
    idx = ctx.index(W.ref, k)
    d_codes = torch.from_numpy(codes).to(dev)
    d_quals = torch.from_numpy(quals).to(dev)
    d_lens = torch.from_numpy(lens).to(dev)
    d_seeds, d_ns = ctx.collect_seeds(idx, d_codes, d_lens, opt=opt_g)
    d_cells = torch.zeros(1, dtype=torch.int64, device=dev)
    d_regs, d_nr = ctx.align_se(idx, d_codes, d_lens, d_cells=d_cells, opt=opt_g)
    torch.cuda.synchronize()
    g_regs_se = d_regs.cpu().numpy().view(_lib.REG_DTYPE).reshape(-1, _lib.MAX_REGS)
    g_nr_se = d_nr.cpu().numpy().copy()
    g_pes = ctx.pestat(idx, d_regs, d_nr, n_pairs, opt=opt_g)
    # the rescue stage on its own (on copies: pair_finish below runs it again as part of the paired stage)
    d_regs_resc, d_nr_resc = d_regs.clone(), d_nr.clone()
    d_stats = torch.zeros(2, dtype=torch.int64, device=dev)
    ctx.mate_rescue(idx, d_codes, d_lens, d_regs_resc, d_nr_resc, resc_pes, d_stats=d_stats, opt=opt_g)
    torch.cuda.synchronize()
    d_alns = ctx.pair_finish(idx, d_codes, d_lens, d_regs, d_nr, g_pes, pair_id0=pair0, opt=opt_g)
    d_counts = torch.zeros(_lib.NCH * idx.l_pac, dtype=torch.int32, device=dev)
    po_g = None
    if popt:
        po_g = _lib.default_pileup_opt()
        for kk, v in popt.items():
            setattr(po_g, kk, v)
    ctx.pileup_accumulate(idx, d_alns, d_codes, d_quals, d_lens, d_counts, popt=po_g)
    rows = ctx.counts_to_rows(idx, d_counts)
    torch.cuda.synchronize()
    g = dict(seeds=d_seeds.cpu().numpy().view(_lib.SEED_DTYPE).reshape(-1, _lib.MAX_SEEDS), n_seeds=d_ns.cpu().numpy(),
             regs_se=g_regs_se, n_regs=g_nr_se, cells=int(d_cells.item()), pes=g_pes,
             regs_resc=d_regs_resc.cpu().numpy().view(_lib.REG_DTYPE).reshape(-1, _lib.MAX_REGS), n_regs_resc=d_nr_resc.cpu().numpy(),
             resc_stats=tuple(int(x) for x in d_stats.cpu()),
             alns=d_alns.cpu().numpy().view(_lib.ALN_DTYPE), counts=rows.cpu().numpy(),
             planes=d_counts.cpu().numpy().reshape(_lib.NCH, idx.l_pac))
    oo = dict(seeds=o["seeds"], n_seeds=o["n_seeds"], regs_se=o_regs_se, n_regs=o_nr_se, cells=o["cells"], pes=o_pes,
              regs_resc=o_regs_resc.reshape(-1, qmo_py.MAX_REGS), n_regs_resc=o_nr_resc, resc_stats=o_resc_stats,
              alns=o_alns, counts=o_counts)
    idx.close()
    return g, oo


def compare(g, o):
    assert np.array_equal(g["n_seeds"], o["n_seeds"])
    for r in range(len(o["n_seeds"])):
        n = o["n_seeds"][r]
        assert np.array_equal(g["seeds"][r, :n], o["seeds"][r, :n]), f"seeds of read {r}"
    assert np.array_equal(g["n_regs"], o["n_regs"])
    bad = [r for r in range(len(o["n_regs"])) if not np.array_equal(g["regs_se"][r, :o["n_regs"][r]], o["regs_se"][r, :o["n_regs"][r]])]
    assert not bad, f"{len(bad)} reads with different regions, first {bad[:5]}"
    assert g["cells"] == o["cells"]
    assert g["pes"].tobytes() == o["pes"].tobytes(), (g["pes"], o["pes"])
    # mate rescue: alignments run, their cells, and every region list afterwards
    assert g["resc_stats"] == o["resc_stats"], (g["resc_stats"], o["resc_stats"])
    assert np.array_equal(g["n_regs_resc"], o["n_regs_resc"])
    bad = [r for r in range(len(o["n_regs_resc"]))
           if not np.array_equal(g["regs_resc"][r, :o["n_regs_resc"][r]], o["regs_resc"][r, :o["n_regs_resc"][r]])]
    assert not bad, f"{len(bad)} reads with different regions after mate rescue, first {bad[:5]}"
    # alignment records: compare field by field; cigar only up to n_cigar
    for f in ("rid", "pos", "flag", "mapq", "n_cigar", "score", "sub", "nm", "mate_rid", "mate_pos", "tlen", "qb", "qe"):
        d = np.nonzero(g["alns"][f] != o["alns"][f])[0]
        assert d.size == 0, f"field {f}: {d.size} records differ, first {d[:5]}: {g['alns'][f][d[:5]]} vs {o['alns'][f][d[:5]]}"
    nc = np.where(o["alns"]["n_cigar"] == 255, 0, o["alns"]["n_cigar"])
    mask = np.arange(g["alns"]["cigar"].shape[1])[None, :] < nc[:, None]
    assert np.array_equal(np.where(mask, g["alns"]["cigar"], 0), np.where(mask, o["alns"]["cigar"], 0))
    assert np.array_equal(g["counts"], o["counts"])
    assert np.array_equal(g["planes"].T, o["counts"])


@pytest.mark.parametrize("name,n", [("cfg1", 3000), ("cfg2", 3000), ("cfg4", 3000), ("cfg5", 2000), ("cfg3", 3000),
                                    ("cfg2-rescue-heavy", 30000), ("cfg5", 12000)])
def test_pipeline_parity(ctx, name, n):
    W = _workload(name, n)
    g, o = run_both(ctx, W, n)
    compare(g, o)
    mapped = (o["alns"]["flag"] & 4) == 0
    assert mapped.mean() > 0.8
    assert o["resc_stats"][0] > 0 and (o["n_regs_resc"] != o["n_regs"]).any()      # rescue ran and placed something
    assert o["counts"][:, 14].sum() > 0


def test_pipeline_without_mate_rescue(ctx):
    """bwa mem -S (QM_F_NO_RESCUE): the paired stage skips mem_matesw on both sides; fewer reads are placed than with it"""
    from quasimodo_b200 import _lib
    W = _workload("cfg1", 3000)
    g, o = run_both(ctx, W, 3000, flags=_lib.F_NO_RESCUE)
    compare(g, o)
    g2, _ = run_both(ctx, W, 3000)
    placed = lambda a: int(((a["flag"] & 4) == 0).sum())
    assert placed(g2["alns"]) > placed(g["alns"])


def test_mate_rescue_all_four_orientations(ctx):
    """the standalone rescue stage with a usable insert-size model in every orientation (FF, FR, RF, RR): mate searched as is and
    reverse-complemented, at the smaller and at the larger coordinate"""
    W = _workload("cfg1", 2500)
    g, o = run_both(ctx, W, 2500, all_orientations=True)
    compare(g, o)
    base = run_both(ctx, W, 2500)[1]["resc_stats"]
    assert o["resc_stats"][0] > 2 * base[0]            # the other three orientations really ran


def test_pipeline_nondefault_scoring(ctx):
    """another scoring scheme (a = 2, unequal gap costs) through every stage, mate rescue included"""
    W = _workload("cfg5", 1500)
    g, o = run_both(ctx, W, 1500, scoring=dict(a=2, b=5, o_del=5, e_del=2, o_ins=7, e_ins=1, T=50, pen_unpaired=25))
    compare(g, o)
    assert o["resc_stats"][0] > 0 and ((o["alns"]["flag"] & 4) == 0).mean() > 0.8


@pytest.mark.parametrize("case", [
    dict(scoring=dict(w=8, zdrop=20, pen_clip5=0, pen_clip3=9)),          # narrow band (retries), early z-drop, unequal end bonuses
    dict(scoring=dict(min_seed_len=20), k=20),                            # -k 20: shorter seeds, more of them
    dict(ragged=True),                                                    # reads of 35 .. 150 bases in one batch
    dict(scoring=dict(T=60, pen_unpaired=5, mask_level=0.3, drop_ratio=0.7)),
    dict(popt=dict(min_bq=25, min_mapq=30, count_orphans=1)),
    dict(popt=dict(ignore_overlaps=1)),
])
def test_pipeline_option_variants(ctx, case):
    """options off their defaults, stage by stage against the oracle"""
    W = _workload("cfg5" if "scoring" in case and "w" in case["scoring"] else "cfg1", 2000)
    g, o = run_both(ctx, W, 2000, **case)
    compare(g, o)
    assert ((o["alns"]["flag"] & 4) == 0).mean() > 0.6 and o["counts"][:, 14].sum() > 0


def test_pipeline_reads_full_of_n(ctx):
    """3 % of the bases are N (30x the simulator's default): N handling of seeding, extension, rescue, CIGAR and pileup"""
    W = _workload("cfg1", 2500)
    W.params.n_ppm = 30000
    g, o = run_both(ctx, W, 2500)
    compare(g, o)
    assert 0.3 < ((o["alns"]["flag"] & 4) == 0).mean() and o["resc_stats"][0] > 100


def test_pipeline_long_reads(ctx):
    """2 x 400 bp: the widest classes of every kernel (extension queries up to 369, rescue with 16 columns per lane, long
    CIGAR tasks, pileup rows of 400 bases)"""
    from quasimodo_b200 import workloads
    W = workloads.Workload("long", [("Merlin", 10), ("TB40E", 3)], ["Merlin"], 800, 77, read_len=400, w=200, indel_ppm=300,
                           ins_mean=900, ins_sd=80, ins_max=1600)
    g, o = run_both(ctx, W, 800)
    compare(g, o)
    assert o["resc_stats"][0] > 0 and ((o["alns"]["flag"] & 4) == 0).mean() > 0.8


def test_pipeline_pair_offset(ctx):
    """a shard starting at pair 5000 gives the records the full run gives for those pairs (index-addressable input)"""
    W = _workload("cfg1", 8000)
    g, o = run_both(ctx, W, 1500, pair0=5000)
    compare(g, o)


def test_pileup_start_channel_counts_admitted_reads(ctx):
    """channel 15 sums to the number of admitted reads (mapped, primary, proper pair)"""
    W = _workload("cfg1", 2000)
    g, o = run_both(ctx, W, 2000)
    a = o["alns"]
    admitted = ((a["flag"] & 4) == 0) & ((a["flag"] & 2) != 0) & (a["n_cigar"] != 0) & (a["n_cigar"] != 255)
    assert int(g["counts"][:, 15].sum()) == int(admitted.sum())


def test_mate_rescue_inline_path_is_bit_exact():
    """QM_RESCUE_INLINE=1 empties the up-front alignment list, so the per-pair pass runs every local alignment itself (the
    path it otherwise takes only for an alignment that becomes necessary after an earlier hit was removed again)"""
    import os
    import subprocess
    import sys
    if os.environ.get("QM_RESCUE_INLINE"):
        pytest.skip("already inside the QM_RESCUE_INLINE run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_pipeline_gpu.py", "-q", "-x", "-k", "cfg1-3000 or cfg5 or without_mate_rescue"],
                       cwd=root, env=dict(os.environ, QM_RESCUE_INLINE="1"), capture_output=True, text=True)
    assert p.returncode == 0 and " passed" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("env", [{"QM_TAIL_MIN": "0"}, {"QM_TAIL_MIN": "0", "QM_SPEC_DEPTH": "2"}, {"QM_SPEC": "0"}],
                         ids=["all-seeds-ahead", "two-seeds-ahead-per-pass", "serial-tail"])
def test_speculative_finish_variants_are_bit_exact(env):
    """the last reads of a batch run every remaining seed's extensions ahead of the state machine (align.cu spec_* kernels);
    a batch as small as a test's goes to the serial warp-per-read tail by default, so QM_TAIL_MIN=0 sends it down the speculative
    path instead; QM_SPEC_DEPTH=2 enters only two seeds per pass (reads run past their directory and open further passes);
    QM_SPEC=0 never speculates.  All give the oracle's regions and cell counts."""
    import os
    import subprocess
    import sys
    if os.environ.get("QM_SPEC_DEPTH") or os.environ.get("QM_SPEC") or os.environ.get("QM_TAIL_MIN"):
        pytest.skip("already inside a QM_SPEC run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_pipeline_gpu.py", "-q", "-x", "-k", "cfg1-3000 or cfg5-2000 or rescue-heavy or long_reads"],
                       cwd=root, env=dict(os.environ, **env), capture_output=True, text=True)
    assert p.returncode == 0 and " passed" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
