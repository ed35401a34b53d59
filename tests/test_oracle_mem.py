"""Oracle sanity (CPU): the bwa-mem restatement must put simulated reads back where the simulator drew them from,
produce self-consistent records, and its pileup must count what the records say.  These are the checks that stand
in for golden vectors the reference does not have for its third-party aligner."""
import numpy as np
import pytest

from oracle import qmo_py
from quasimodo_b200 import workloads


@pytest.fixture(scope="module")
def run():
    n = 3000
    W = workloads.config1(n)
    codes, quals, src, pos = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, cells, pes = qmo_py.run_sample(ref, codes, quals, lens)
    return dict(W=W, n=n, codes=codes, quals=quals, src=src, pos=pos, alns=alns, counts=counts, cells=cells, pes=pes)


def test_reference_strain_reads_map_to_origin(run):
    s, strand = run["src"] & 0xffff, run["src"] >> 16
    start, ins = run["pos"] & ((1 << 40) - 1), run["pos"] >> 40
    merlin = s == 1                                    # source 1 = Merlin = the alignment reference
    a0, a1 = run["alns"][0::2], run["alns"][1::2]
    left = np.where(strand == 0, a0["pos"], a1["pos"])
    both = ((a0["flag"] & 4) == 0) & ((a1["flag"] & 4) == 0)
    ok = (left == start) & both
    assert ok[merlin].mean() > 0.97
    # proper pairs with TLEN = insert size
    tl = np.abs(a0["tlen"])
    assert (tl[merlin & ok] == ins[merlin & ok]).mean() > 0.97
    assert ((a0["flag"] & 2) != 0)[merlin].mean() > 0.97


def test_insert_size_model(run):
    fr = run["pes"][1]
    assert not fr["failed"] and abs(fr["avg"] - 350) < 3 and abs(fr["std"] - 35) < 3
    assert all(run["pes"][d]["failed"] for d in (0, 2, 3))


def test_records_self_consistent(run):
    a = run["alns"]
    mapped = (a["flag"] & 4) == 0
    assert mapped.mean() > 0.9
    for r in a[mapped][:500]:
        ql = sum(c >> 4 for c in r["cigar"][:r["n_cigar"]] if (c & 0xf) in (0, 1, 4))
        assert ql == 150
        assert r["qe"] - r["qb"] == sum(c >> 4 for c in r["cigar"][:r["n_cigar"]] if (c & 0xf) in (0, 1))
        assert 0 <= r["mapq"] <= 60 and r["pos"] >= 0
    # mates point at each other
    a0, a1 = a[0::2], a[1::2]
    both = ((a0["flag"] & 4) == 0) & ((a1["flag"] & 4) == 0)
    assert np.array_equal(a0["mate_pos"][both], a1["pos"][both]) and np.array_equal(a1["mate_pos"][both], a0["pos"][both])
    assert np.array_equal(a0["tlen"][both], -a1["tlen"][both])


def test_pileup_totals(run):
    a, cnt = run["alns"], run["counts"]
    admitted = ((a["flag"] & 4) == 0) & ((a["flag"] & 2) != 0) & (a["n_cigar"] != 0) & (a["n_cigar"] != 255)
    m_bases = sum(int(sum(c >> 4 for c in r["cigar"][:r["n_cigar"]] if (c & 0xf) == 0)) for r in a[admitted])
    assert int(cnt[:, 14].sum()) == m_bases
    assert int(cnt[:, 15].sum()) == int(admitted.sum())
    assert (cnt[:, :5].sum() + cnt[:, 6:11].sum()) <= m_bases          # BQ filter and overlap zeroing only remove
    assert (cnt >= 0).all()


def test_shards_sum_to_whole(run):
    """oracle statement of the multi-GPU decomposition: shard counts (each with the sample's prefix) add up"""
    from quasimodo_b200 import sharding
    W, n = run["W"], run["n"]
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    lens = np.full(2 * n, 150, np.int32)
    total = np.zeros_like(run["counts"])
    for r in range(3):
        lo, hi = sharding.shard_range(n, r, 3)
        pre = None if r == 0 and not sharding.needs_prefix(lo, hi, n, qmo_py.PESTAT_PAIRS) else (run["codes"], lens)
        _, c, _, pes = qmo_py.run_sample(ref, run["codes"][2 * lo:2 * hi], run["quals"][2 * lo:2 * hi], lens[2 * lo:2 * hi],
                                         pair_id0=lo, prefix=pre)
        assert pes.tobytes() == run["pes"].tobytes()
        total += c
    assert np.array_equal(total, run["counts"])


def test_packed_genomes_equal_bwas_own_index_files():
    """the genomes this repo ships (quasimodo_b200/data/genomes/*.qmg) against digests of the reference's own bwa index
    (ref/*.pac + *.ann, tests/golden/make_pac_digests.py): same contigs, same lengths, same bases"""
    import hashlib
    import json
    import os
    from quasimodo_b200 import genomes
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pac_digests.json")))
    assert set(gold) == {"Merlin", "TB40E", "AD169", "Phix", "Ecoli"}
    for stem, g in gold.items():
        G = genomes.load(stem)
        assert G.names == g["names"] and G.lens == g["lens"] and G.total == g["l_pac"], stem
        assert hashlib.sha256(G.codes.tobytes()).hexdigest() == g["sha256_codes"], stem


def test_text_pileup_agrees_with_the_count_tensor(run):
    """the two restatements of the column walk (counts, samtools-mpileup text) agree column by column"""
    import re
    text = qmo_py.mpileup_text(qmo_py.Ref(run["W"].ref.codes, run["W"].ref.lens, k=31), run["alns"], run["codes"], run["quals"],
                               np.full(2 * run["n"], 150, np.int32), ["Merlin"])
    cnt = run["counts"]
    seen = np.zeros(len(cnt), bool)
    n_zero = 0
    for line in text.split(b"\n"):
        if not line:
            continue
        chrom, pos, refb, depth, bases, quals = line.split(b"\t")
        p = int(pos) - 1
        seen[p] = True
        if int(depth) == 0:          # every base of the column failed -Q: samtools prints "*" for both strings
            assert bases == b"*" and quals == b"*" and int(cnt[p, 0:5].sum() + cnt[p, 6:11].sum()) == 0
            n_zero += 1
            continue
        assert chrom == b"Merlin" and refb == b"ACGT"[run["W"].ref.codes[p]:run["W"].ref.codes[p] + 1]
        assert int(depth) == len(quals)
        # drop ^X, $ and indel strings, then one character per entry
        out, i = [], 0
        while i < len(bases):
            c = bases[i:i + 1]
            if c == b"^":
                i += 2
            elif c == b"$":
                i += 1
            elif c in b"+-":
                m = re.match(rb"\d+", bases[i + 1:])
                i += 1 + len(m.group()) + int(m.group())
            else:
                out.append(c)
                i += 1
        assert len(out) == int(depth)
        n_star = sum(1 for c in out if c == b"*")
        assert int(depth) - n_star == int(cnt[p, 0:5].sum() + cnt[p, 6:11].sum())
        assert n_star <= int(cnt[p, 5] + cnt[p, 11])
        fwd_ref = sum(1 for c in out if c == b".")
        rev_ref = sum(1 for c in out if c == b",")
        r = run["W"].ref.codes[p]
        assert fwd_ref == cnt[p, r] and rev_ref == cnt[p, 6 + r]
    # a column has a line iff an admitted read covers it (raw depth or a deletion)
    assert np.array_equal(seen, (cnt[:, 14] + cnt[:, 5] + cnt[:, 11]) > 0)
    assert n_zero > 0


def test_mem_patch_reg_has_nothing_to_patch_on_baseline_data():
    """the product has no mem_patch_reg (bwamem.c: two collinear hits of a read become one when a banded global
    alignment explains them at >= 90 % of the predicted score).  The oracle restates the rule behind QMO_F_PATCH; on reads of the
    BASELINE configs it finds nothing to do: short reads' collinear hit pairs are rare and fail the relative-bandwidth test, so the
    regions with and without the flag are identical."""
    import ctypes as C
    from quasimodo_b200 import workloads
    L = qmo_py.lib()
    tried0 = C.c_longlong.in_dll(L, "g_qmo_patch_tried").value
    for W, rl, w in ((workloads.config2(4, 6000), 150, 100), (workloads.config2(1, 6000), 150, 100), (workloads.config5(4000), 250, 200)):
        n = 6000 if rl == 150 else 4000
        codes, _, _, _ = W.simulate_host(0, n)
        lens = np.full(2 * n, rl, np.int32)
        ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
        o0, o1 = qmo_py.default_opt(), qmo_py.default_opt()
        o0.w = o1.w = w
        o1.flags |= 4                                  # QMO_F_PATCH
        a0 = qmo_py.align_se(ref, codes, lens, opt=o0)
        a1 = qmo_py.align_se(ref, codes, lens, opt=o1)
        assert np.array_equal(a0["n_regs"], a1["n_regs"]) and a0["regs"].tobytes() == a1["regs"].tobytes()
        assert (a0["n_regs"] > 1).sum() > 50           # there ARE reads with several hits: repeats, other strands -- not collinear pairs
    done = C.c_longlong.in_dll(L, "g_qmo_patch_done").value
    assert done == 0 and C.c_longlong.in_dll(L, "g_qmo_patch_tried").value - tried0 <= 5


@pytest.mark.parametrize("cfg, n, popt", [
    ("cfg1", 1200, {}),                                                          # 2 x 150, substitutions only, mates overlap often
    ("cfg5", 600, {}),                                                           # 2 x 250 with simulated indels
    ("cfg3", 800, {}),                                                           # three contigs, unmapped contaminant reads
    ("cfg1", 600, dict(min_bq=25, min_mapq=30)),
    ("cfg5", 400, dict(count_orphans=1, ignore_overlaps=1)),
])
def test_count_tensor_equals_an_independent_restatement(cfg, n, popt):
    """bcftools is not in the image, so the counting oracle cannot be pinned; SURVEY.md Appendix A asks for two independent
    restatements diffed against each other instead.  oracle/pileup_py.py goes htslib's way (pileup entries per read, the second
    mate looked up in a hash of the first), oracle/qmo_pileup.c walks the pair's CIGARs with two cursors: same tensor, every channel."""
    from oracle import pileup_py
    W = {"cfg1": workloads.config1, "cfg5": workloads.config5, "cfg3": workloads.config3}[cfg](n)
    codes, quals, _, _ = W.simulate_host(0, n)
    L = W.params.read_len
    lens = np.full(2 * n, L, np.int32)
    lens[[3, 10]] = [L - 30, L // 2]                                             # ragged reads
    codes[3, L - 30:] = 4
    codes[10, L // 2:] = 4
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    alns = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)[0]
    po = qmo_py.PileupOpt(0, 13, 0, 0)
    for k, v in popt.items():
        setattr(po, k, v)
    want = qmo_py.pileup(ref, alns, codes, quals, lens, po)
    offs = np.concatenate([[0], np.cumsum(W.ref.lens)])
    got = pileup_py.count_tensor(offs, ref.l_pac, alns, codes, quals, lens, po.min_mapq, po.min_bq, bool(po.count_orphans), bool(po.ignore_overlaps))
    assert np.array_equal(got, want), np.argwhere(got != want)[:5]
    assert want[:, 14].sum() > n * L and (cfg != "cfg5" or (want[:, 12].sum() > 0 and want[:, 13].sum() > 0 and want[:, 5].sum() > 0))


@pytest.mark.parametrize("cfg, n", [("cfg1", 3000), ("cfg5", 1500), ("cfg3", 2500)])
def test_insert_size_model_equals_an_independent_restatement(cfg, n):
    """mem_pestat twice: oracle/qmo_mem.c against oracle/pestat_py.py (written separately from SURVEY.md A.6) on the same
    region lists -- windows, failed flags and the double-precision mean / deviation, bit for bit"""
    from oracle import pestat_py
    W = {"cfg1": workloads.config1, "cfg5": workloads.config5, "cfg3": workloads.config3}[cfg](n)
    codes, _, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    o = qmo_py.align_se(ref, codes, lens, opt=opt)
    want = qmo_py.pestat(ref, o["regs"], o["n_regs"], opt=opt)
    got, n_obs = pestat_py.pestat(ref.l_pac, o["regs"], o["n_regs"])
    assert n_obs[1] > 0.8 * n * (0.85 if cfg == "cfg3" else 1)          # FR pairs dominate
    for d in range(4):
        assert bool(want[d]["failed"]) == bool(got[d]["failed"]), d
        if not got[d]["failed"]:
            assert (int(want[d]["low"]), int(want[d]["high"])) == (got[d]["low"], got[d]["high"]), d
            assert float(want[d]["avg"]) == got[d]["avg"] and float(want[d]["std"]) == got[d]["std"], d
    assert not got[1]["failed"]


def test_insert_size_model_all_orientations_synthetic():
    """hand-made region lists: all four orientation bins filled (one below the 5 % rule, one below ten samples), outliers,
    second-best regions that do and do not disqualify a pair, mates on different contigs"""
    from oracle import pestat_py
    W = workloads.config3(10)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    l_pac = ref.l_pac
    rng = np.random.default_rng(9)
    n = 6000
    regs = np.zeros((2 * n, qmo_py.MAX_REGS), dtype=qmo_py.REG_DTYPE)
    n_regs = np.ones(2 * n, dtype=np.int32)
    for p in range(n):
        u = rng.random()
        d = 1 if u < 0.80 else 0 if u < 0.93 else 3 if u < 0.998 else 2       # FR 80 %, FF 13 %, RR 6.8 % , RF ~0.2 % (< 10 samples)
        ins = int(rng.normal(300, 30)) if rng.random() > 0.03 else int(rng.integers(1, 12000))
        b1 = int(rng.integers(20000, 200000))
        if rng.random() < 0.5:
            b1 = 2 * l_pac - 1 - b1                                               # read 1 on the reverse strand
        r1 = b1 >= l_pac
        # invert mem_infer_dir: pick p2 on read 1's strand, then map it back to read 2's own strand
        same = d in (0, 3)
        p2 = b1 + ins if d in (0, 1) else b1 - ins
        b2 = p2 if same else 2 * l_pac - 1 - p2
        for e, b in ((0, b1), (1, b2)):
            r = regs[2 * p + e]
            r[0]["rb"], r[0]["re"], r[0]["qb"], r[0]["qe"], r[0]["score"], r[0]["rid"] = b, b + 150, 0, 150, 140, 0
            if rng.random() < 0.2:                                                # a second region: overlapping or not, strong or weak
                n_regs[2 * p + e] = 2
                qb = int(rng.choice([0, 100]))
                r[1]["rb"], r[1]["re"], r[1]["qb"], r[1]["qe"], r[1]["rid"] = b + 5000, b + 5100, qb, qb + 50 + int(rng.integers(0, 60)), 0
                r[1]["score"] = int(rng.choice([40, 111, 112, 113, 130]))
        if rng.random() < 0.02:
            regs[2 * p + 1][0]["rid"] = 1
        if rng.random() < 0.02:
            n_regs[2 * p] = 0
    want = qmo_py.pestat(ref, regs, n_regs)
    got, n_obs = pestat_py.pestat(l_pac, regs, n_regs)
    assert min(n_obs[0], n_obs[1], n_obs[3]) > 200 and 0 < n_obs[2]
    assert [bool(x["failed"]) for x in got] == [False, False, True, False] or n_obs[2] >= 10
    for d in range(4):
        assert bool(want[d]["failed"]) == bool(got[d]["failed"]), (d, n_obs)
        if not got[d]["failed"]:
            assert (int(want[d]["low"]), int(want[d]["high"])) == (got[d]["low"], got[d]["high"]), d
            assert float(want[d]["avg"]) == got[d]["avg"] and float(want[d]["std"]) == got[d]["std"], d


@pytest.mark.parametrize("cfg, n, depth", [("cfg1", 3000, 3), ("cfg1", 3000, 8), ("cfg3", 2500, 2), ("cfg5", 1500, 5), ("cfg1", 1500, 250)])
def test_depth_cap_equals_a_literal_replay_of_the_iterator(cfg, n, depth):
    """`bcftools mpileup -d N`: which reads htslib's pileup iterator keeps is order dependent; oracle/qmo_pileup.c restates it with a
    heap of read ends, oracle/depthcap_py.py replays the linked list literally -- the same reads survive"""
    from oracle import depthcap_py
    W = {"cfg1": workloads.config1, "cfg5": workloads.config5, "cfg3": workloads.config3}[cfg](n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    alns = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)[0]
    want = qmo_py.depth_cap(ref, alns, depth)
    got = depthcap_py.depth_cap(alns, depth)
    assert np.array_equal(got, want), (int(got.sum()), int(want.sum()), np.flatnonzero(got != want)[:10])
    everyone = qmo_py.depth_cap(ref, alns, 1 << 30)
    assert (want <= everyone).all() and (depth >= 250 or want.sum() < everyone.sum())


def test_indel_allele_table_equals_an_independent_restatement():
    """config 5 (simulated indels): every allele (anchor, type, length, inserted bases) and its forward / reverse support, from
    oracle/qmo_pileup.c and from oracle/pileup_py.py"""
    from oracle import pileup_py
    n = 1500
    W = workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    codes[7, 40:44] = 4                                                          # Ns, possibly inside an insertion
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    alns = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)[0]
    rows = qmo_py.indels(ref, alns, codes, lens)
    want = {(int(r[0]), int(r[1]), int(r[3]), int(r[2]), int(r[5]), int(r[4])): [int(r[6]), int(r[7])] for r in rows}
    got = pileup_py.indel_alleles(np.concatenate([[0], np.cumsum(W.ref.lens)]), alns, codes, lens)
    assert got == want
    assert len(want) > 100 and any(k[2] == 0 for k in want) and any(k[2] == 1 for k in want) and any(sum(v) > 1 for v in want.values())


def test_hash_seeds_against_brute_force():
    """the default seeder's contract: on every diagonal, the maximal runs of read positions whose k-mer equals the
    reference's there, counting only k-mers that occur 1..32 times on that strand and lie inside one contig.  Brute force over a
    small two-contig genome with a 3-copy repeat and a 40-copy repeat (above the cap), reads from both strands with errors."""
    rng = np.random.default_rng(17)
    k, cap = 31, 32
    unit200, unit40 = rng.integers(0, 4, 200), rng.integers(0, 4, 40)
    c0 = np.concatenate([rng.integers(0, 4, 900), unit200, rng.integers(0, 4, 500), unit200, rng.integers(0, 4, 300)] +
                        [np.concatenate([unit40, rng.integers(0, 4, 3)]) for _ in range(40)])
    c1 = np.concatenate([rng.integers(0, 4, 700), unit200, rng.integers(0, 4, 600)])
    fwd = np.concatenate([c0, c1]).astype(np.uint8)
    lens_ref = np.array([len(c0), len(c1)], np.int64)
    l_pac = len(fwd)
    D = np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])
    # contig interval of every doubled position (mirrored for the reverse half): a k-mer must not leave it
    bounds = [(0, len(c0)), (len(c0), l_pac), (l_pac, l_pac + len(c1)), (l_pac + len(c1), 2 * l_pac)]
    where = {}
    for lo, hi in bounds:
        for x in range(lo, hi - k + 1):
            where.setdefault((lo >= l_pac, D[x:x + k].tobytes()), []).append(x)
    n = 60
    L = 150
    reads = np.full((n, L), 4, np.uint8)
    for r in range(n):
        x = int(rng.integers(0, 2 * l_pac - L))
        rd = D[x:x + L].copy()
        for j in np.flatnonzero(rng.random(L) < 0.02):
            rd[j] = (rd[j] + 1 + rng.integers(0, 3)) % 4
        if r % 7 == 0:
            rd[int(rng.integers(0, L))] = 4
        reads[r] = rd
    ref = qmo_py.Ref(fwd, lens_ref, k=k)
    o = qmo_py.align_se(ref, reads, np.full(n, L, np.int32))
    n_multi = 0
    for r in range(n):
        hits = set()
        for q in range(L - k + 1):
            km = reads[r, q:q + k]
            if (km > 3).any():
                continue
            for half in (False, True):
                occ = where.get((half, km.tobytes()), [])
                if 1 <= len(occ) <= cap:
                    hits.update((q, x) for x in occ)
        want = []
        for q, x in sorted(hits):
            if (q - 1, x - 1) in hits:
                continue                                                    # not the start of a run
            run = 1
            while (q + run, x + run) in hits:
                run += 1
            want.append((q, x, k + run - 1))
        got = sorted((int(s["qbeg"]), int(s["rbeg"]), int(s["len"])) for s in o["seeds"][r][:o["n_seeds"][r]])
        assert len(want) < qmo_py.MAX_SEEDS
        assert got == sorted(want), (r, got, want)
        n_multi += len(want) > 1
    assert n_multi > 10


@pytest.mark.parametrize("cfg, n, popt", [("cfg1", 800, {}), ("cfg5", 500, {}), ("cfg3", 700, {}), ("cfg5", 300, dict(min_bq=30, ignore_overlaps=1))])
def test_text_pileup_equals_an_independent_restatement(cfg, n, popt):
    """samtools-mpileup text twice (oracle/qmo_pileup.c and oracle/pileup_py.py): byte-identical, indel strings, read starts and
    ends, deleted bases and all-filtered columns included"""
    from oracle import pileup_py
    W = {"cfg1": workloads.config1, "cfg5": workloads.config5, "cfg3": workloads.config3}[cfg](n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    alns = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)[0]
    po = qmo_py.PileupOpt(0, 13, 0, 0)
    for k, v in popt.items():
        setattr(po, k, v)
    names = [f"contig{i}" for i in range(len(W.ref.lens))]
    want = qmo_py.mpileup_text(ref, alns, codes, quals, lens, names, po)
    offs = np.concatenate([[0], np.cumsum(W.ref.lens)])
    got = pileup_py.mpileup_text(W.ref.codes, offs, W.ref.lens, names, alns, codes, quals, lens, po.min_mapq, po.min_bq,
                                 bool(po.count_orphans), bool(po.ignore_overlaps))
    if got != want:
        gl, wl = got.split(b"\n"), want.split(b"\n")
        bad = next(i for i in range(min(len(gl), len(wl))) if gl[i] != wl[i])
        raise AssertionError((len(gl), len(wl), gl[bad][:300], wl[bad][:300]))
    assert want.count(b"\n") > 10000 and (cfg != "cfg5" or (b"+1" in want and b"-1" in want and b"*" in want))


@pytest.mark.parametrize("cfg", ["cfg2", "cfg5"])
def test_single_end_finish_equals_an_independent_restatement(cfg):
    """batches of 8 pairs have no insert-size model (mem_pestat wants 10 per orientation): no rescue, no pairing, every record is
    the read's best hit with mem_approx_mapq_se's quality.  Order of the hits (score, then bwa's hash of the read id), primary,
    sub-optimal score, rivals and MAPQ from oracle/mapq_py.py against the records of oracle/qmo_mem.c -- on a sample whose
    references share long repeats, so that ties and sub-optimal hits are common"""
    from oracle import mapq_py
    n, step = (1600, 8) if cfg == "cfg2" else (800, 8)
    W = workloads.config2(4, n) if cfg == "cfg2" else workloads.config5(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    se = qmo_py.align_se(ref, codes, lens, opt=opt)
    l_pac = ref.l_pac
    seen = dict(mapped=0, unmapped=0, with_sub=0, mapq0=0, mapq60=0, mid=0)
    for p0 in range(0, n, step):
        sl = slice(2 * p0, 2 * (p0 + step))
        alns, _, _, pes = qmo_py.run_sample(ref, codes[sl], quals[sl], lens[sl], pair_id0=p0, opt=opt)
        assert all(pes[d]["failed"] for d in range(4))
        for j in range(2 * step):
            r = 2 * p0 + j
            want = mapq_py.finish_single_end(se["regs"][r], int(se["n_regs"][r]), r)
            a = alns[j]
            if want is None:
                assert a["flag"] & 4, r
                seen["unmapped"] += 1
                continue
            assert not (a["flag"] & 4) or a["n_cigar"] == 255, r
            assert (int(a["score"]), int(a["sub"]), int(a["mapq"])) == (want["score"], want["sub"], want["mapq"]), (r, a, want)
            assert bool(a["flag"] & 0x10) == (want["rb"] >= l_pac), r
            seen["mapped"] += 1
            seen["with_sub"] += want["sub"] > 0
            seen["mapq0"] += want["mapq"] == 0
            seen["mapq60"] += want["mapq"] == 60
            seen["mid"] += 0 < want["mapq"] < 60
    assert seen["mapped"] > 0.9 * 2 * n and seen["with_sub"] > 30 and seen["mapq0"] > 20 and seen["mid"] > 5, seen


@pytest.mark.parametrize("cfg", ["cfg2", "cfg5", "cfg3"])
def test_pairing_decision_equals_an_independent_restatement(cfg):
    """mem_pair + the MAPQ logic of mem_sam_pe twice (mate rescue off, so both see the same hit lists): which hits are reported, the
    proper-pair flag, AS / XS and the mapping quality of every record"""
    from oracle import pair_py
    n = 1500
    W = {"cfg2": lambda: workloads.config2(4, n), "cfg5": lambda: workloads.config5(n), "cfg3": lambda: workloads.config3(n)}[cfg]()
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    opt.flags = qmo_py.F_NO_RESCUE
    se = qmo_py.align_se(ref, codes, lens, opt=opt)
    pes = qmo_py.pestat(ref, se["regs"], se["n_regs"], opt=opt)
    alns = qmo_py.pair_and_finish(ref, codes, lens, se["regs"].copy(), se["n_regs"].copy(), pes, opt=opt)
    offs = np.concatenate([[0], np.cumsum(W.ref.lens)])
    seen = dict(proper=0, improper=0, unmapped=0, lifted=0)
    for p in range(n):
        proper, recs = pair_py.finish_pair(ref.l_pac, offs, pes, se["regs"][2 * p], int(se["n_regs"][2 * p]), se["regs"][2 * p + 1],
                                           int(se["n_regs"][2 * p + 1]), p)
        for m in (0, 1):
            a, want = alns[2 * p + m], recs[m]
            if want is None:
                assert a["flag"] & 4, (p, m)
                seen["unmapped"] += 1
                continue
            if a["n_cigar"] == 255:
                continue                                   # CIGAR too long for the record: reported unmapped by both implementations
            assert not (a["flag"] & 4), (p, m)
            assert (int(a["score"]), int(a["sub"]), int(a["mapq"]), bool(a["flag"] & 0x10)) == \
                (want["score"], want["sub"], want["mapq"], want["rb"] >= ref.l_pac), (p, m, a, want)
            assert bool(a["flag"] & 2) == bool(proper), (p, m)
        seen["proper" if proper else "improper"] += 1
        if proper and recs[0] and recs[1]:
            hs = pair_py.order_and_mark(se["regs"][2 * p], int(se["n_regs"][2 * p]), 2 * p)
            seen["lifted"] += recs[0]["mapq"] > pair_py.approx_mapq(hs[0]) if hs and hs[0]["rb"] == recs[0]["rb"] else 0
    assert seen["proper"] > 0.8 * n * (0.85 if cfg == "cfg3" else 1) and seen["improper"] > 3 and seen["lifted"] > 0, seen


@pytest.mark.parametrize("cfg", ["cfg5", "cfg2"])
def test_record_generation_equals_an_independent_restatement(cfg):
    """mem_reg2aln + bwa_gen_cigar2 twice: the hit the single-end finish reports (batches without an insert-size model, so the hit is
    known) turned into position, strand, CIGAR and NM by oracle/cigar_py.py over oracle/ksw_py.py's whole-matrix global alignment,
    against the records of oracle/qmo_mem.c.  Config 5 has indels (gapped paths, leftmost placement on the forward strand)."""
    from oracle import cigar_py, mapq_py
    n, step = (320, 8) if cfg == "cfg5" else (480, 8)
    W = workloads.config5(n) if cfg == "cfg5" else workloads.config2(4, n)
    codes, quals, _, _ = W.simulate_host(0, n)
    L = W.params.read_len
    lens = np.full(2 * n, L, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    se = qmo_py.align_se(ref, codes, lens, opt=opt)
    fwd = np.asarray(W.ref.codes, dtype=np.uint8)
    doubled = np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])
    offs = [int(x) for x in np.concatenate([[0], np.cumsum(W.ref.lens)])]
    n_gapped = n_clipped = n_rev_gapped = 0
    for p0 in range(0, n, step):
        sl = slice(2 * p0, 2 * (p0 + step))
        alns = qmo_py.run_sample(ref, codes[sl], quals[sl], lens[sl], pair_id0=p0, opt=opt)[0]
        for j in range(2 * step):
            r = 2 * p0 + j
            fin = mapq_py.finish_single_end(se["regs"][r], int(se["n_regs"][r]), r)
            a = alns[j]
            if fin is None or a["n_cigar"] == 255:
                continue
            rec = cigar_py.hit_to_record(doubled, ref.l_pac, offs, codes[r, :L], fin["hit"], w=W.w)
            got = [(int(c) & 15, int(c) >> 4) for c in a["cigar"][:int(a["n_cigar"])]]
            assert (int(a["rid"]), int(a["pos"]), bool(a["flag"] & 0x10), got, int(a["nm"])) == \
                (rec["rid"], rec["pos"], rec["rev"], rec["cigar"], rec["nm"]), (r, a, rec)
            gapped = any(op in (1, 2) for op, _ in got)
            n_gapped += gapped
            n_rev_gapped += gapped and rec["rev"]
            n_clipped += any(op == 4 for op, _ in got)
    assert n_clipped > 5 and (cfg != "cfg5" or (n_gapped > 20 and n_rev_gapped > 5)), (n_gapped, n_rev_gapped, n_clipped)


@pytest.mark.parametrize("cfg", ["cfg1", "cfg3"])
def test_flags_and_mate_fields_follow_the_sam_rules(cfg):
    """every pair of records against the SAM-level rules bwa's mem_aln2sam implements, stated from the SAM side: 0x1 / 0x40 / 0x80, 0x8
    and 0x20 mirror the mate's 0x4 and 0x10, an unmapped read sits at its mate's coordinate (and takes its strand bit), RNEXT / PNEXT
    are the mate's, TLEN spans the two 5' ends (signed, +-1 inclusive, 0 when a mate is unmapped or on another contig), 0x2 only
    when both are mapped to one contig"""
    n = 2500
    W = {"cfg1": workloads.config1, "cfg3": workloads.config3}[cfg](n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    lens[[4, 9]] = [20, 25]                                                   # too short to seed: unmapped reads with mapped mates
    codes[4, 20:] = 4
    codes[9, 25:] = 4
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns = qmo_py.run_sample(ref, codes, quals, lens)[0]
    seen = dict(both=0, one=0, none=0, cross=0)
    for p in range(n):
        pair = alns[2 * p:2 * p + 2]
        placed = [not (int(x["flag"]) & 4) for x in pair]
        ends5 = []
        for x in pair:
            span = sum(int(c) >> 4 for c in x["cigar"][:int(x["n_cigar"])] if (int(c) & 15) in (0, 2)) if int(x["n_cigar"]) != 255 else 0
            ends5.append(int(x["pos"]) + (span - 1 if int(x["flag"]) & 0x10 else 0))
        for m in (0, 1):
            me, mate, f = pair[m], pair[1 - m], int(pair[m]["flag"])
            assert f & 0x1 and bool(f & 0x40) == (m == 0) and bool(f & 0x80) == (m == 1) and not (f & 0x900)
            assert bool(f & 0x8) == (not placed[1 - m])
            if placed[m] or placed[1 - m]:
                assert bool(f & 0x20) == bool(int(mate["flag"]) & 0x10)
                assert (int(me["mate_rid"]), int(me["mate_pos"])) == (int(mate["rid"]), int(mate["pos"]))
            else:
                assert (int(me["rid"]), int(me["mate_rid"]), int(me["tlen"])) == (-1, -1, 0)
            if not placed[m] and placed[1 - m]:
                assert (int(me["rid"]), int(me["pos"])) == (int(mate["rid"]), int(mate["pos"])) and bool(f & 0x10) == bool(int(mate["flag"]) & 0x10)
            if placed[0] and placed[1] and int(pair[0]["rid"]) == int(pair[1]["rid"]):
                d = ends5[1 - m] - ends5[m]
                assert int(me["tlen"]) == d + (d > 0) - (d < 0), (p, m)
            else:
                assert int(me["tlen"]) == 0 and not (f & 0x2)
        seen["both" if all(placed) else "none" if not any(placed) else "one"] += 1
        seen["cross"] += all(placed) and int(pair[0]["rid"]) != int(pair[1]["rid"])
    assert seen["both"] > 0.85 * n and seen["one"] >= 2 and (cfg != "cfg3" or seen["none"] + seen["cross"] >= 0), seen


@pytest.mark.parametrize("cfg", ["cfg2", "cfg5"])
def test_mate_rescue_equals_an_independent_restatement(cfg):
    """mem_matesw and its loop twice: which windows are searched (count and cells of the local alignments), what each find becomes,
    how the mate's hit list ends up (order, redundancy filter) -- oracle/rescue_py.py against oracle/qmo_mem.c on a sample where
    one strain's reads often lack a hit of their own (TB40E reads against AD169)"""
    from oracle import rescue_py
    n = 1200 if cfg == "cfg2" else 500
    W = workloads.config2(1, n) if cfg == "cfg2" else workloads.config5(n)
    codes, _, _, _ = W.simulate_host(0, n)
    L = W.params.read_len
    lens = np.full(2 * n, L, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = W.w
    se = qmo_py.align_se(ref, codes, lens, opt=opt)
    pes = qmo_py.pestat(ref, se["regs"], se["n_regs"], opt=opt)
    regs, n_regs = se["regs"].copy(), se["n_regs"].copy()
    want_sw, want_cells = qmo_py.mate_rescue(ref, codes, lens, regs, n_regs, pes, opt=opt)
    fwd = np.asarray(W.ref.codes, dtype=np.uint8)
    doubled = np.concatenate([fwd, (3 - fwd[::-1]).astype(np.uint8)])
    offs = [int(x) for x in np.concatenate([[0], np.cumsum(W.ref.lens)])]
    fields = ("rid", "score", "csub", "qb", "qe", "rb", "re")
    counters = dict(sw=0, cells=0)
    n_new = 0
    for p in range(n):
        hits = [[{f: int(r[f]) for f in fields} for r in se["regs"][2 * p + m][:int(se["n_regs"][2 * p + m])]] for m in (0, 1)]
        got = rescue_py.rescue_pair(doubled, ref.l_pac, offs, W.ref.lens, pes, hits, [codes[2 * p, :L], codes[2 * p + 1, :L]], opt, counters)
        for m in (0, 1):
            want = [{f: int(r[f]) for f in fields} for r in regs[2 * p + m][:int(n_regs[2 * p + m])]]
            assert got[m] == want, (p, m, got[m], want)
            n_new += len(want) > len(hits[m])
    assert (counters["sw"], counters["cells"]) == (want_sw, want_cells) and want_sw > (50 if cfg == "cfg2" else 5) and n_new > (20 if cfg == "cfg2" else 2), (counters, want_sw, n_new)
