"""GPU tests of the file-level drop-in: the device radix sort against a stable argsort, and the C++
driver end to end -- FASTQ in; sorted BAM + BAI, count TSV, VCF and cleaned FASTQ out -- against the CPU oracle's records
ordered by the restated samtools comparator."""
import ctypes as C
import gzip
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("n,bits", [(0, 24), (1, 24), (2, 8), (4095, 24), (4096, 13), (4097, 38), (100_003, 24), ((1 << 20) + 3, 64), (3_000_000, 30)])
def test_radix_sort_is_a_stable_sort(ctx, n, bits):
    from quasimodo_b200 import _lib
    rng = np.random.default_rng(n + bits)
    hi = (1 << bits) - 1
    keys = rng.integers(0, hi, n, dtype=np.uint64, endpoint=True)
    if n > 10:                                       # long runs of equal keys: stability is the point
        keys[rng.integers(0, n, n // 2)] = keys[0]
        keys[: n // 4] = keys[: n // 4] & np.uint64(0xff)
    perm = np.full(n, 0xffffffff, dtype=np.uint32)
    rc = _lib.lib().qm_sort_keys_host(ctx._h, keys.ctypes.data, n, bits, perm.ctypes.data)
    assert rc == 0, _lib.lib().qm_last_error(ctx._h)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))


def test_device_keys_match_header_formula(ctx):
    import torch
    from oracle import qmo_py, sort_py
    from quasimodo_b200 import _lib, workloads
    W = workloads.Workload("t", [("Merlin", 1), ("Phix", 1)], ["Merlin", "Phix"], 10, 1)
    idx = ctx.index(W.ref, 31)
    rng = np.random.default_rng(5)
    n = 50_000
    a = np.zeros(n, dtype=qmo_py.ALN_DTYPE)
    a["rid"] = rng.integers(-1, 2, n)
    a["pos"] = np.where(a["rid"] >= 0, rng.integers(0, 5000, n), -1)
    a["flag"] = rng.integers(0, 2, n) * 0x10
    d_alns = torch.from_numpy(a.view(np.uint8)).cuda()
    d_keys = torch.empty(n, dtype=torch.int64, device="cuda")
    d_perm = torch.empty(n, dtype=torch.int32, device="cuda")
    bits = C.c_int()
    L = _lib.lib()
    assert L.qm_aln_sort_keys(ctx._h, idx._h, C.c_void_p(d_alns.data_ptr()), n, C.c_void_p(d_keys.data_ptr()), C.byref(bits), None) == 0
    assert L.qm_sort_pairs(ctx._h, C.c_void_p(d_keys.data_ptr()), C.c_void_p(d_perm.data_ptr()), n, bits.value, None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_perm.cpu().numpy().view(np.uint32), sort_py.sort_perm(a))
    k = d_keys.cpu().numpy()
    assert (np.diff(k) >= 0).all()
    idx.close()


@pytest.fixture(scope="module")
def sample_case(tmp_path_factory):
    """3000 pairs of a 4-source mixture against Merlin|Phix (E. coli reads stay unmapped): oracle records + driver outputs"""
    from oracle import qmo_py, sort_py
    from quasimodo_b200 import workloads
    from tests import bamio, drvutil
    d = tmp_path_factory.mktemp("drvgpu")
    n = 3000
    me = 235000
    W = workloads.Workload("t", [("AD169", 1), ("Merlin", 10), ("Phix", 1), ("Ecoli", 1)], ["Merlin", "Phix"], n, 78,
                           extra_weights=[40 * me, 40 * me, 10 * me, 10 * me])
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    for r, l in ((3, 120), (8, 75), (2001, 149)):
        lens[r] = l
        codes[r, l:] = 4
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, counts, cells, _ = qmo_py.run_sample(ref, codes, quals, lens)
    names = drvutil.pair_names("sim.g", n)
    fa, r1, r2 = str(d / "ref.fa"), str(d / "r1.fq.gz"), str(d / "r2.fq.gz")
    drvutil.write_fasta(W.ref, fa)
    drvutil.write_fastq(codes, quals, lens, names, r1, r2, gz=True)
    bam, tsv, vcf, txt = str(d / "s.bam"), str(d / "s.mpileup"), str(d / "s.vcf"), str(d / "s.text.mpileup")
    p = drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--bam", bam, "--counts", tsv, "--vcf", vcf, "--sample", "TM-1-1",
                            "--min-dp", 3, "--min-alt", 2, "-t", 4, "--mpileup", txt])
    return dict(W=W, n=n, codes=codes, quals=quals, lens=lens, alns=alns, counts=counts, cells=cells, names=names, perm=sort_py.sort_perm(alns),
                bam=bamio.Bam(bam), bam_path=bam, tsv=tsv, vcf=vcf, txt=txt, ref=ref, fa=fa, r1=r1, r2=r2, dir=d, stderr=p.stderr)


def test_driver_bam_matches_oracle_records_in_samtools_order(sample_case):
    from tests import bamio, drvutil
    assert f"{sample_case['n']} pairs aligned, {sample_case['cells']} extension cells" in sample_case["stderr"]
    assert drvutil.check_bam_records(sample_case) > sample_case["n"]
    refs, n_no_coor = bamio.read_bai(sample_case["bam_path"] + ".bai")
    assert n_no_coor == int((sample_case["alns"]["rid"] < 0).sum()) > 0


def test_driver_text_pileup_matches_oracle(sample_case):
    """--mpileup: the samtools-mpileup text of the sample (reads of ragged lengths, two contigs) equals the oracle's, byte for byte"""
    from oracle import qmo_py
    c = sample_case
    want = qmo_py.mpileup_text(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], list(c["W"].ref.names))
    got = open(c["txt"], "rb").read()
    assert len(want) > 100000 and got == want


def test_driver_baq_caps_counts_and_text_pileup(sample_case):
    """--baq 1 (both mpileups of the reference flow run without -B): count TSV and text pileup equal the oracle's on the qualities its
    BAQ restatement caps; without the flag the outputs above are the -B ones"""
    from oracle import qmo_py
    from tests import drvutil
    c = sample_case
    d = c["dir"]
    tsv, txt = str(d / "baq.tsv"), str(d / "baq.text.mpileup")
    drvutil.run_driver(["sample", "--ref", c["fa"], "--r1", c["r1"], "--r2", c["r2"], "--counts", tsv, "--mpileup", txt, "--baq", 1, "-t", 2])
    capped = qmo_py.baq(c["ref"], c["alns"], c["codes"], c["quals"], c["lens"], flag=3)
    want = qmo_py.pileup(c["ref"], c["alns"], c["codes"], capped, c["lens"])
    assert not np.array_equal(want, c["counts"])
    got = np.array([[int(x) for x in ln.rstrip("\n").split("\t")[4:]] for ln in list(open(tsv))[1:]], dtype=np.int32)
    assert np.array_equal(got, want)
    want_txt = qmo_py.mpileup_text(c["ref"], c["alns"], c["codes"], capped, c["lens"], list(c["W"].ref.names))
    assert open(txt, "rb").read() == want_txt != open(c["txt"], "rb").read()


def test_driver_count_tsv_matches_oracle(sample_case):
    W, counts = sample_case["W"], sample_case["counts"]
    rows = [ln.rstrip("\n").split("\t") for ln in open(sample_case["tsv"])]
    assert rows[0][:4] == ["chrom", "pos", "ref", "depth"] and len(rows) == 1 + W.ref.total
    body = rows[1:]
    got = np.array([[int(x) for x in r[4:]] for r in body], dtype=np.int32)
    assert np.array_equal(got, counts)
    assert [r[0] for r in body[:3]] == [W.ref.names[0]] * 3 and body[-1][0] == W.ref.names[-1]
    assert body[0][1] == "1" and body[W.ref.lens[0]][1] == "1" and body[W.ref.lens[0] - 1][1] == str(W.ref.lens[0])
    assert "".join(r[2] for r in body[:50]) == "".join("ACGT"[c] for c in W.ref.codes[:50])
    depth = counts[:, 0:5].sum(1) + counts[:, 6:11].sum(1)
    assert [int(r[3]) for r in body] == depth.tolist()


def test_driver_vcf_matches_python_writer(ctx, sample_case, tmp_path):
    from quasimodo_b200 import _lib, formats
    W = sample_case["W"]
    idx = ctx.index(W.ref, 31)
    s = ctx.sample(idx)
    s.add_pairs_host(sample_case["codes"], sample_case["quals"], sample_case["lens"])
    copt = _lib.default_call_opt()
    copt.min_dp, copt.min_alt = 3, 2
    calls = s.call_snps(copt)
    assert len(calls) > 20
    want = str(tmp_path / "py.vcf")
    formats.write_vcf(want, W.ref, "TM-1-1", calls)
    a = [ln for ln in open(want) if not ln.startswith("##reference")]
    full = [ln for ln in open(sample_case["vcf"]) if not ln.startswith("##reference")]
    indel_hdr = ("##INFO=<ID=INDEL", "##INFO=<ID=IDV", "##INFO=<ID=ADF", "##INFO=<ID=ADR")
    b = [ln for ln in full if not ln.startswith(indel_hdr) and "\tINDEL;" not in ln]
    assert a == b
    # the indel records (AD169 reads against Merlin carry real strain indels): anchor base first, sorted in with the SNPs
    ind = [ln.split("\t") for ln in full if "\tINDEL;" in ln]
    assert len(ind) > 3
    codes, names = W.ref.codes, list(W.ref.names)
    offs = np.concatenate([[0], np.cumsum(W.ref.lens)])
    for f in ind:
        o = offs[names.index(f[0])] + int(f[1]) - 1
        assert f[3][0] == f[4][0] == "ACGT"[codes[o]] and (len(f[3]) == 1) != (len(f[4]) == 1)
        assert f[3] == "".join("ACGT"[c] for c in codes[o:o + len(f[3])])
    pos = [(names.index(ln.split("\t")[0]), int(ln.split("\t")[1])) for ln in full if not ln.startswith("#")]
    assert pos == sorted(pos)
    s.close()
    idx.close()


def test_driver_batches_do_not_change_the_outputs(tmp_path):
    """70,000 pairs in batches of 65,536 (2 batches) and in one batch: identical records and counts"""
    from quasimodo_b200 import workloads
    from tests import bamio, drvutil
    n = 70_000
    W = workloads.config1(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    fa, r1, r2 = str(tmp_path / "ref.fa"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    drvutil.write_fastq(codes, quals, lens, drvutil.pair_names("p", n), r1, r2)
    outs = []
    for tag, bp in (("a", 65536), ("b", 1 << 20)):
        bam, tsv = str(tmp_path / f"{tag}.bam"), str(tmp_path / f"{tag}.tsv")
        drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--bam", bam, "--counts", tsv, "--batch-pairs", bp])
        data = b"".join(d for _, d in bamio.bgzf_blocks(bam))
        l_text = int.from_bytes(data[4:8], "little")
        outs.append((data[8 + l_text:], open(tsv, "rb").read()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert len(outs[0][0]) > 2 * n * 200


def test_driver_scoring_options_reach_the_aligner(tmp_path):
    """bwa mem's -A -B -O -E -L -U -T -d on the driver's command line: the records in the BAM are the oracle's under the same options
    (the scheme of test_pipeline_nondefault_scoring; -A 2 alone would also double -d and -L, so they are given)"""
    from oracle import qmo_py, sort_py
    from quasimodo_b200 import workloads
    from tests import bamio, drvutil
    n = 1200
    W = workloads.config1(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    opt = qmo_py.default_opt()
    for name, v in dict(a=2, b=5, o_del=5, e_del=2, o_ins=7, e_ins=1, T=50, pen_unpaired=25).items():
        setattr(opt, name, v)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, _, cells, _ = qmo_py.run_sample(ref, codes, quals, lens, opt=opt)
    base = qmo_py.run_sample(ref, codes, quals, lens)[0]
    assert (alns["score"] != base["score"]).mean() > 0.5            # the options matter
    names = drvutil.pair_names("sc", n)
    fa, r1, r2, bam = str(tmp_path / "ref.fa"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq"), str(tmp_path / "s.bam")
    drvutil.write_fasta(W.ref, fa)
    drvutil.write_fastq(codes, quals, lens, names, r1, r2)
    p = drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--bam", bam, "-A", 2, "-B", 5, "-O", "5,7", "-E", "2,1", "-T", 50,
                            "-U", 25, "-d", 100, "-L", 5])
    assert f"{n} pairs aligned, {cells} extension cells" in p.stderr
    case = dict(W=W, codes=codes, quals=quals, lens=lens, alns=alns, names=names, perm=sort_py.sort_perm(alns), bam=bamio.Bam(bam))
    assert drvutil.check_bam_records(case) > n


def test_driver_decontam(sample_case):
    """rules/decontamination.smk:15-17: pairs with both records unmapped against the contaminant survive, in input order"""
    from oracle import qmo_py
    from quasimodo_b200 import genomes
    from tests import drvutil
    d = sample_case["dir"]
    phix = genomes.load("Phix")
    pfa = str(d / "phix.fa")
    drvutil.write_fasta(phix, pfa)
    o1, o2 = str(d / "clean1.fq"), str(d / "clean2.fq")
    drvutil.run_driver(["decontam", "--ref", pfa, "--r1", sample_case["r1"], "--r2", sample_case["r2"], "--out-r1", o1, "--out-r2", o2])
    ref = qmo_py.Ref(phix.codes, phix.lens, k=31)
    alns, _, _, _ = qmo_py.run_sample(ref, sample_case["codes"], sample_case["quals"], sample_case["lens"])
    f = alns["flag"].reshape(-1, 2)
    keep = np.flatnonzero((((f & 12) == 12) & ((f & 256) == 0)).all(1))
    assert 0 < len(keep) < sample_case["n"]

    def expect(m):
        out = []
        for p in keep:
            r = 2 * p + m
            L = sample_case["lens"][r]
            out += [f"@{sample_case['names'][p]}/{m + 1}", "".join("ACGTN"[c] for c in sample_case["codes"][r, :L]), "+",
                    "".join(chr(q + 33) for q in sample_case["quals"][r, :L])]
        return out
    assert open(o1).read().split("\n")[:-1] == expect(0)
    assert open(o2).read().split("\n")[:-1] == expect(1)
    # fused pass: Merlin is the target (first contig), PhiX the contaminant -> drop pairs with a mate mapped on PhiX
    drvutil.run_driver(["decontam", "--ref", sample_case["fa"], "--keep-contigs", 1, "--r1", sample_case["r1"], "--r2", sample_case["r2"],
                        "--out-r1", o1, "--out-r2", o2])
    a = sample_case["alns"].reshape(-1, 2)
    on_phix = (((a["flag"] & 4) == 0) & (a["rid"] >= 1)).any(1)
    got = [ln[1:-3] for ln in open(o1) if ln.startswith("@sim.g.")]
    assert got == [sample_case["names"][p] for p in np.flatnonzero(~on_phix)]


def test_driver_fails_loudly_without_outputs(tmp_path, sample_case):
    from tests import drvutil
    p = drvutil.run_driver(["sample", "--ref", sample_case["fa"], "--r1", sample_case["r1"]], check=False)
    assert p.returncode == 1
    p = drvutil.run_driver(["sample", "--ref", sample_case["fa"], "--r1", sample_case["r1"], "--r2", sample_case["r2"], "--gpu", 99], check=False)
    assert p.returncode == 3 and "no usable B200" in p.stderr


def test_driver_edge_inputs(tmp_path):
    """empty input, a single pair, reads shorter than the seed length, all-N reads: the driver still writes valid files"""
    from quasimodo_b200 import genomes
    from tests import bamio, drvutil
    phix = genomes.load("Phix")
    fa = str(tmp_path / "phix.fa")
    drvutil.write_fasta(phix, fa)
    seq = "".join("ACGT"[c] for c in phix.codes[100:250])
    rc = "".join("ACGT"[3 - c] for c in phix.codes[300:450][::-1])
    cases = {
        "empty": [],
        "one": [("p0", seq, rc)],
        "odd": [("p0", seq, rc), ("short", seq[:20], rc[:25]), ("allN", "N" * 150, "N" * 150), ("p1", seq[:100], rc[:131])],
    }
    for tag, pairs in cases.items():
        r1, r2 = str(tmp_path / f"{tag}.r1.fq"), str(tmp_path / f"{tag}.r2.fq")
        with open(r1, "w") as f1, open(r2, "w") as f2:
            for nm, a, b in pairs:
                f1.write(f"@{nm}/1\n{a}\n+\n{'I' * len(a)}\n")
                f2.write(f"@{nm}/2\n{b}\n+\n{'I' * len(b)}\n")
        bam, tsv, vcf = (str(tmp_path / f"{tag}.{x}") for x in ("bam", "tsv", "vcf"))
        drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--bam", bam, "--counts", tsv, "--vcf", vcf])
        B = bamio.Bam(bam)
        assert len(B.records) == 2 * len(pairs)
        assert sum(1 for _ in open(tsv)) == 1 + phix.total
        assert open(vcf).read().startswith("##fileformat=VCFv4.2")
        bamio.read_bai(bam + ".bai")
        if tag != "empty":
            first = [r for r in B.records if r["name"] == "p0"]
            assert sorted(r["pos"] for r in first) == [100, 300] and all(r["tags"]["NM"] == 0 for r in first)
        if tag == "odd":
            by = {(r["name"], r["flag"] & 0x40): r for r in B.records}
            assert by[("short", 0x40)]["flag"] & 4 and by[("allN", 0x40)]["flag"] & 4 and by[("allN", 0)]["seq"] == "N" * 150
            assert by[("p1", 0x40)]["pos"] == 100 and by[("p1", 0)]["cigar"] == [131 << 4]


def test_driver_writes_every_declared_output_of_the_replaced_rules(sample_case, tmp_path):
    """one `qm_driver sample` call at the reference's own path patterns (quasimodo_b200/rules.py, checked against the rule files
    in tests/test_rules_cpu.py): sorted BAM + BAI, duplicate-free BAM + BAI, metrics, text pileup, VCF, bgzipped VCF + tabix index"""
    import gzip
    from quasimodo_b200 import rules
    from tests import bamio, drvutil
    dirs = {k: str(tmp_path / k) for k in ("seq_dir", "snpcall_dir", "report_dir")}
    argv, paths = rules.sample_command(drvutil.driver_path(), dirs, "TM-1-1", "Merlin", sample_case["fa"], sample_case["r1"], sample_case["r2"],
                                       extra=["--min-dp", "3", "--min-alt", "2"])
    drvutil.run_driver(argv[1:])
    for p in paths:
        assert os.path.exists(p) and os.path.getsize(p) > 0, p
    vcf = rules.expand(dirs, "bcftools", "vcf", "TM-1-1", "Merlin")
    assert gzip.open(vcf + ".gz", "rb").read() == open(vcf, "rb").read()
    tbi = bamio.read_tbi(vcf + ".gz.tbi")
    text = bamio.BgzfText(vcf + ".gz")
    recs = [ln.rstrip("\n").encode() for ln in open(vcf) if not ln.startswith("#")]
    assert len(recs) > 20 and tbi["names"] == [recs[0].split(b"\t")[0].decode()]
    name = tbi["names"][0]
    assert bamio.tabix_query(tbi, text, name, 0, 1 << 29) == recs
    lo, hi = int(recs[5].split(b"\t")[1]) - 1, int(recs[9].split(b"\t")[1])
    assert bamio.tabix_query(tbi, text, name, lo, hi) == recs[5:10]
    kept = bamio.Bam(rules.expand(dirs, "rmdup", "rmdupbam", "TM-1-1", "Merlin"))
    assert 0 < len(kept.records) <= 2 * sample_case["n"] and not any(r["flag"] & 0x400 for r in kept.records)
    bamio.read_bai(rules.expand(dirs, "rmdup", "rmdupbam", "TM-1-1", "Merlin") + ".bai")


def test_driver_on_two_gpus_writes_the_same_files(tmp_path):
    """--gpus 0,1: batches dealt over two GPUs, the first batch's insert-size model handed to both, counts merged with one NCCL
    all-reduce -- BAM records, count TSV and VCF identical to the single-GPU run"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from quasimodo_b200 import workloads
    from tests import bamio, drvutil
    n = 300_000
    W = workloads.config1(n)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    fa, r1, r2 = str(tmp_path / "ref.fa"), str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    drvutil.write_fastq(codes, quals, lens, drvutil.pair_names("p", n), r1, r2)
    outs = []
    for tag, gp in (("one", ["--gpu", 0]), ("two", ["--gpus", "0,1"])):
        bam, tsv, vcf = (str(tmp_path / f"{tag}.{x}") for x in ("bam", "tsv", "vcf"))
        p = drvutil.run_driver(["sample", "--ref", fa, "--r1", r1, "--r2", r2, "--bam", bam, "--counts", tsv, "--vcf", vcf, "--batch-pairs", 65536] + gp)
        data = b"".join(d for _, d in bamio.bgzf_blocks(bam))
        l_text = int.from_bytes(data[4:8], "little")
        outs.append((data[8 + l_text:], open(tsv, "rb").read(), [ln for ln in open(vcf) if not ln.startswith("##reference")], p.stderr))
    assert "count tensors of 2 GPUs merged" in outs[1][3]
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2]
