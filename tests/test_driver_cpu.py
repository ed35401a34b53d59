"""CPU tests of the C++ host driver's file formats: the BAM + BAI writer is run on alignment
records produced by the CPU oracle (`qm_driver bam-from-records --perm`, which touches no GPU) and read back with the
test-side BAM reader; record order is checked against the restated samtools comparator (oracle/sort_py.py)."""
import re

import numpy as np
import pytest

from oracle import qmo_py, sort_py
from quasimodo_b200 import workloads
from tests import bamio, drvutil

N_PAIRS = 500


@pytest.fixture(scope="module")
def case(tmp_path_factory):
    d = tmp_path_factory.mktemp("drv")
    ad, me = 229000, 235000
    W = workloads.Workload("t", [("AD169", 1), ("Merlin", 10), ("Phix", 1), ("Ecoli", 1)], ["Merlin", "Phix"], N_PAIRS, 77,
                           extra_weights=[40 * ad, 40 * me, 10 * me, 10 * me])
    codes, quals, _, _ = W.simulate_host(0, N_PAIRS)
    lens = np.full(2 * N_PAIRS, 150, np.int32)
    lens[5] = 97                                    # ragged input
    lens[10] = 60
    codes[5, 97:] = 4
    codes[10, 60:] = 4
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns, _, _, _ = qmo_py.run_sample(ref, codes, quals, lens)
    names = drvutil.pair_names("sim.t", N_PAIRS)
    fa, r1, r2 = str(d / "ref.fa"), str(d / "r1.fq.gz"), str(d / "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    drvutil.write_fastq(codes, quals, lens, names, r1, str(d / "r2.tmp"), gz=True, comment=" 1:N:0")
    drvutil.write_fastq(codes, quals, lens, names, str(d / "r1.tmp"), r2, gz=False)
    perm = sort_py.sort_perm(alns)
    alns.tofile(str(d / "alns.bin"))
    perm.tofile(str(d / "perm.bin"))
    bam = str(d / "out.bam")
    drvutil.run_driver(["bam-from-records", "--ref", fa, "--r1", r1, "--r2", r2, "--alns", d / "alns.bin", "--perm", d / "perm.bin",
                        "--bam", bam, "-t", 3])
    return dict(W=W, codes=codes, quals=quals, lens=lens, alns=alns, names=names, perm=perm, bam=bamio.Bam(bam), bam_path=bam,
                dir=d, fa=fa, r1=r1, r2=r2)


def test_compact_key_orders_like_samtools():
    """the library's compact sort key (include/quasimodo_b200.h: qm_sort_key) must order records exactly like the
    literal samtools key, ties included"""
    rng = np.random.default_rng(3)
    n, n_contigs, max_len = 20000, 3, 5000
    a = np.zeros(n, dtype=qmo_py.ALN_DTYPE)
    a["rid"] = rng.integers(-1, n_contigs, n)
    a["pos"] = np.where(a["rid"] >= 0, rng.integers(0, max_len, n), -1)
    a["flag"] = rng.integers(0, 2, n) * 0x10
    pos_bits = 1
    while (1 << pos_bits) <= max_len + 1:
        pos_bits += 1
    rid = np.where(a["rid"] < 0, n_contigs, a["rid"]).astype(np.uint64)
    compact = (rid << np.uint64(pos_bits + 1)) | ((a["pos"].astype(np.int64) + 1).astype(np.uint64) << np.uint64(1)) | \
              ((a["flag"] & 0x10) != 0).astype(np.uint64)
    assert np.array_equal(np.argsort(compact, kind="stable"), sort_py.sort_perm(a))
    assert a["rid"][sort_py.sort_perm(a)][-1] == -1                  # unplaced records last


def test_bam_header_and_order(case):
    bam, W = case["bam"], case["W"]
    assert bam.refs == list(zip(W.ref.names, W.ref.lens))
    assert bam.text.startswith("@HD\tVN:1.6\tSO:coordinate\n")
    assert [ln.split("\t")[1][3:] for ln in bam.text.split("\n") if ln.startswith("@SQ")] == W.ref.names
    assert len(bam.records) == 2 * N_PAIRS
    keys = [((r["rid"] & 0xffffffff) << 32) | ((r["pos"] + 1) << 1) | ((r["flag"] >> 4) & 1) for r in bam.records]
    assert keys == sorted(keys)
    assert any(r["rid"] == -1 for r in bam.records) and any(r["flag"] & 0x10 for r in bam.records)


def test_bam_records_match_oracle(case):
    assert drvutil.check_bam_records(case) > N_PAIRS


def test_bai_finds_every_overlapping_record(case):
    bam = case["bam"]
    refs, n_no_coor = bamio.read_bai(case["bam_path"] + ".bai")
    assert len(refs) == len(bam.refs)
    assert n_no_coor == sum(1 for r in bam.records if r["rid"] == -1)
    rng = np.random.default_rng(11)
    spans = []
    for r in bam.records:
        rlen = sum(x >> 4 for x in r["cigar"] if (x & 15) in (0, 2))
        spans.append((r["rid"], r["pos"], r["pos"] + max(rlen, 1)))
    for rid, (name, ln) in enumerate(bam.refs):
        bins, lin = refs[rid]
        recs = [i for i, s in enumerate(spans) if s[0] == rid]
        if not recs:
            assert not bins
            continue
        meta = bins.pop(37450)
        assert meta[1] == (sum(1 for i in recs if not bam.records[i]["flag"] & 4), sum(1 for i in recs if bam.records[i]["flag"] & 4))
        # every record lies in exactly the chunk list of its own bin
        for i in recs:
            u = bam.records[i]["ustart"]
            assert any(bam.voffset_to_u(b) <= u < bam.voffset_to_u(e) for b, e in bins[bam.records[i]["bin"]]), i
        # linear index: the first record overlapping each 16 kb window
        for w, v in enumerate(lin):
            ov = [i for i in recs if spans[i][1] < (w + 1) << 14 and spans[i][2] > w << 14]
            if ov:
                assert bam.voffset_to_u(v) == bam.records[ov[0]]["ustart"], (rid, w)
        # region queries through bins + linear index
        for _ in range(20):
            beg = int(rng.integers(0, ln))
            end = min(ln, beg + int(rng.integers(1, 40000)))
            want = {i for i in recs if spans[i][1] < end and spans[i][2] > beg}
            lo = bam.voffset_to_u(lin[beg >> 14]) if (beg >> 14) < len(lin) else None
            got = set()
            for b in sort_py.reg2bins(beg, end):
                for cb, ce in bins.get(b, []):
                    ub, ue = bam.voffset_to_u(cb), bam.voffset_to_u(ce)
                    got |= {i for i in recs if ub <= bam.records[i]["ustart"] < ue and (lo is None or bam.records[i]["uend"] > lo)}
            assert want <= got, (rid, beg, end)


def test_driver_rejects_bad_input(case):
    d = case["dir"]
    p = drvutil.run_driver(["bam-from-records", "--ref", case["fa"], "--r1", case["r1"], "--r2", str(d / "nope.fq"), "--alns", d / "alns.bin",
                            "--perm", d / "perm.bin", "--bam", d / "x.bam"], check=False)
    assert p.returncode == 2 and "cannot open" in p.stderr
    short = str(d / "short.fq")
    with open(case["r2"]) as fh, open(short, "w") as out:
        out.writelines(fh.readlines()[:40])
    p = drvutil.run_driver(["bam-from-records", "--ref", case["fa"], "--r1", case["r1"], "--r2", short, "--alns", d / "alns.bin",
                            "--perm", d / "perm.bin", "--bam", d / "x.bam"], check=False)
    assert p.returncode == 2 and "fewer records" in p.stderr
    bad = str(d / "bad.fa")
    open(bad, "w").write(">c1\nACGTNNACGT\n")
    p = drvutil.run_driver(["bam-from-records", "--ref", bad, "--r1", case["r1"], "--r2", case["r2"], "--alns", d / "alns.bin",
                            "--perm", d / "perm.bin", "--bam", d / "x.bam"], check=False)
    assert p.returncode == 2 and "only A/C/G/T" in p.stderr
    for text, msg in ((">c1\nACGT\n>c2\n>c3\nAC\n", "contig c2 is empty"), (">c1 x\nACGT\n>c1 y\nAC\n", "occurs twice")):
        open(bad, "w").write(text)
        p = drvutil.run_driver(["bam-from-records", "--ref", bad, "--r1", case["r1"], "--r2", case["r2"], "--alns", d / "alns.bin",
                                "--perm", d / "perm.bin", "--bam", d / "x.bam"], check=False)
        assert p.returncode == 2 and msg in p.stderr, p.stderr
    p = drvutil.run_driver(["frobnicate"], check=False)
    assert p.returncode == 1
    # an option the command does not know is a usage error, not a silent run with the default
    p = drvutil.run_driver(["sample", "--ref", case["fa"], "--r1", case["r1"], "--r2", case["r2"], "--min-qual", "3"], check=False)
    assert p.returncode == 1 and "unknown option '--min-qual'" in p.stderr
    for spec, msg in (("0,0", "listed twice"), ("1-0", "no device given"), ("0,-3", "names no device list")):
        p = drvutil.run_driver(["sample", "--ref", case["fa"], "--r1", case["r1"], "--r2", case["r2"], "--gpus", spec], check=False)
        assert p.returncode == 1 and msg in p.stderr, (spec, p.stderr)
    p = drvutil.run_driver(["vcf-index", "--vcf", d / "x.vcf", "--gpu", "0"], check=False)
    assert p.returncode == 1 and "unknown option" in p.stderr


def test_driver_takes_bwa_mems_scoring_options():
    """-A -B -O -E -L -U -T -d -c -D -W mean what they mean to `bwa mem` (rules/bwa.smk:15 passes only -k 31, the defaults below);
    bwa's fastmap.c scales -T -d -B -O -E -L -U by -A unless they are given themselves"""
    def resolved(*args):
        p = drvutil.run_driver(["sample", "--print-options", "1", *args])
        return dict(kv.split("=") for kv in p.stdout.split())
    assert resolved() == dict(q="0", Q="13", **{"count-orphans": "0", "ignore-overlaps": "0"}, A="1", B="4", O="6,6", E="1,1", L="5,5", U="17", T="30", d="100", w="100", k="31", c="500", D="0.5", W="0", flags="0")
    r = resolved("-A", "2")
    assert (r["B"], r["O"], r["E"], r["L"], r["U"], r["T"], r["d"]) == ("8", "12,12", "2,2", "10,10", "34", "60", "200")
    r = resolved("-A", "2", "-B", "5", "-O", "5,7", "-E", "2", "-L", "0,9", "-T", "50", "-U", "25", "-d", "20", "-w", "8", "-c", "50", "-D", "0.3", "-W", "3")
    assert r == dict(q="0", Q="13", **{"count-orphans": "0", "ignore-overlaps": "0"}, A="2", B="5", O="5,7", E="2,2", L="0,9", U="25", T="50", d="20", w="8", k="31", c="50", D="0.3", W="3", flags="0")
    assert resolved("--no-rescue", "1")["flags"] == "1" and resolved("--fm-seeds", "1")["flags"] == "2"
    r = resolved("-q", "20", "-Q", "25", "--count-orphans", "1", "--ignore-overlaps", "1")       # the mpileups' -q -Q -A -x
    assert (r["q"], r["Q"], r["count-orphans"], r["ignore-overlaps"]) == ("20", "25", "1", "1")
    assert resolved("--min-mapq", "7", "--min-bq", "0")["q"] == "7"
    for bad in (["-Q", "94"], ["-O", "6x"], ["-O", "1,2,3"], ["-B", "-4"], ["-A", "0"], ["-w", "0"], ["-k", "32"], ["-D", "1.5"], ["-E", ""]):
        p = drvutil.run_driver(["sample", "--print-options", "1", *bad], check=False)
        assert p.returncode == 1 and "option -" in p.stderr, bad


def test_vcf_gz_and_tabix_index(tmp_path):
    """rule `bcftools` declares {sample}.vcf.gz and tabix-indexes it (rules/vcfcall.smk:107,118-119); `qm_driver vcf-index` is the
    host-only entry of the same writer the sample command uses: the .gz must inflate to the VCF, the .tbi must find exactly
    the records a scan finds, for every kind of region"""
    import gzip
    from tests import bamio, drvutil
    rng = np.random.default_rng(8)
    vcf = tmp_path / "s.vcf"
    rows = []
    for chrom, n, L in (("Merlin", 4000, 235000), ("phiX", 30, 5386), ("ecoli", 2500, 4_641_652)):
        pos = np.sort(rng.choice(np.arange(1, L), n, replace=False))
        for p in pos:
            ref = "ACGT"[rng.integers(0, 4)] * (1 if rng.random() < 0.9 else int(rng.integers(2, 40)))
            rows.append((chrom, int(p), ref))
    with open(vcf, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##contig=<ID=Merlin,length=235000>\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts\n")
        for chrom, p, ref in rows:
            fh.write(f"{chrom}\t{p}\t.\t{ref}\tG\t50\tPASS\tDP=30\tGT\t1\n")
    drvutil.run_driver(["vcf-index", "--vcf", vcf])
    gz = str(vcf) + ".gz"
    assert gzip.open(gz, "rb").read() == open(vcf, "rb").read()
    tbi = bamio.read_tbi(gz + ".tbi")
    assert (tbi["format"], tbi["col_seq"], tbi["col_beg"], tbi["col_end"], tbi["meta"], tbi["skip"]) == (2, 1, 2, 0, ord("#"), 0)
    assert tbi["names"] == ["Merlin", "phiX", "ecoli"]
    for (bins, lin), n in zip(tbi["refs"], (4000, 30, 2500)):
        assert bins[37450][1][0] == n                      # htslib's pseudo-bin: records of the sequence
    text = bamio.BgzfText(gz)
    regions = [("Merlin", 0, 235000), ("Merlin", 1000, 1001), ("Merlin", 16383, 16385), ("Merlin", 100000, 163840), ("phiX", 0, 100),
               ("phiX", 5000, 5386), ("ecoli", 4_000_000, 4_641_652), ("ecoli", 131071, 131073), ("nope", 0, 10)]
    regions += [("ecoli", int(b), int(b) + int(rng.integers(1, 300000))) for b in rng.integers(0, 4_600_000, 20)]
    for name, b, e in regions:
        want = [f"{c}\t{p}\t.\t{r}\tG\t50\tPASS\tDP=30\tGT\t1".encode() for c, p, r in rows if c == name and p - 1 < e and p - 1 + len(r) > b]
        assert bamio.tabix_query(tbi, text, name, b, e) == want, (name, b, e)
    # unsorted input cannot be indexed: non-zero exit, as tabix
    bad = tmp_path / "bad.vcf"
    bad.write_text("#CHROM\tPOS\tID\tREF\tALT\nMerlin\t500\t.\tA\tG\nMerlin\t20\t.\tA\tG\n")
    assert drvutil.run_driver(["vcf-index", "--vcf", bad], check=False).returncode == 2


def test_fastq_side_parses_what_bwa_would(tmp_path):
    """`qm_driver fastq-check` (host only) runs the FASTQ side of `sample`: plain and gzip input, CRLF line ends, blank lines between
    records, a last line without a newline, records longer than the file buffer's remainder; and it fails loudly (exit 2) on what
    bwa would reject: a truncated record, a quality line of another length, mate files of different length."""
    import gzip
    rng = np.random.default_rng(3)
    n = 30_000                                              # ~5 MB per file: several refills of the 4 MB buffer
    seqs = ["".join("ACGTN"[c] for c in rng.integers(0, 5, int(l))) for l in rng.integers(30, 151, n)]
    def records(mate, eol="\n", blank_every=0):
        out = []
        for i, s in enumerate(seqs):
            out.append(f"@r{i}/{mate} extra words{eol}{s}{eol}+{eol}{'I' * len(s)}{eol}")
            if blank_every and i % blank_every == 0:
                out.append(eol)
        return "".join(out)
    p1, p2 = tmp_path / "a_1.fq", tmp_path / "a_2.fq.gz"
    p1.write_text(records(1, "\r\n", blank_every=997).rstrip("\r\n"))          # CRLF, blank lines, no final newline
    with gzip.open(p2, "wt") as fh:
        fh.write(records(2))
    ok = drvutil.run_driver(["fastq-check", "--r1", p1, "--r2", p2])
    total = 2 * sum(len(s) for s in seqs)
    assert ok.stdout.startswith(f"{n} pairs, {total} bases, longest read {max(len(s) for s in seqs)},")
    # malformed inputs
    bad = tmp_path / "bad.fq"
    good = records(1)
    cases = {"truncated": good[:good.rindex("+")], "lengths": good.replace("I" * len(seqs[5]) + "\n", "I" * (len(seqs[5]) + 1) + "\n", 1) if len(seqs[5]) != len(seqs[4]) else None,
             "header": good.replace("@r7/1", "r7/1", 1)}
    for name, text in cases.items():
        if text is None:
            continue
        bad.write_text(text)
        p = drvutil.run_driver(["fastq-check", "--r1", bad, "--r2", p2], check=False)
        assert p.returncode == 2 and str(bad) in p.stderr, (name, p.returncode, p.stderr[-200:])
    short = tmp_path / "short.fq"
    short.write_text("".join(records(1).split("@r29999/1")[:1]))
    p = drvutil.run_driver(["fastq-check", "--r1", short, "--r2", p2], check=False)
    assert p.returncode == 2 and "more records" in p.stderr
    # mate files out of step (a record lost in one of them): bwa's "paired reads have different names", not a silent mis-pairing
    shifted = tmp_path / "shifted.fq"
    shifted.write_text(records(1).replace(f"@r100/1 extra words\n{seqs[100]}\n+\n{'I' * len(seqs[100])}\n", "", 1) + f"@tail/1\nACGT\n+\nIIII\n")
    p = drvutil.run_driver(["fastq-check", "--r1", shifted, "--r2", p2], check=False)
    assert p.returncode == 2 and 'paired reads have different names: "r101"' in p.stderr and '"r100"' in p.stderr, p.stderr[-300:]


def test_driver_selftest():
    """host-only checks of helpers that otherwise only run with several GPUs (the by-key merge of the per-GPU indel allele tables)"""
    assert drvutil.run_driver(["selftest"]).stdout.strip() == "selftest ok"


def test_driver_benchmark_file_has_snakemakes_shape(tmp_path):
    """--benchmark FILE: one header line + one row in the column order of a Snakemake `benchmark:` file, which is what
    scripts/resource_benchmark.R reads (read_tsv(..., skip = 1) with nine column names; it uses s, max_rss and io_out)"""
    vcf = tmp_path / "x.vcf"
    vcf.write_text("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\nc1\t5\t.\tA\tC\t50\t.\tDP=20\n")
    bench = tmp_path / "x.benchmark.txt"
    drvutil.run_driver(["vcf-index", "--vcf", vcf, "--benchmark", bench])
    head, row = bench.read_text().splitlines()
    assert head.split("\t") == ["s", "h:m:s", "max_rss", "max_vms", "max_uss", "max_pss", "io_in", "io_out", "mean_load"]
    f = row.split("\t")
    assert len(f) == 9 and re.fullmatch(r"\d+:\d\d:\d\d", f[1])
    v = [float(x) for x in f[:1] + f[2:]]
    assert 0 <= v[0] < 60 and 1 < v[1] < 4096 and v[2] >= v[1] and v[5] > 0 and all(x >= 0 for x in v) and v[7] < 100 * 64
    assert (tmp_path / "x.vcf.gz.tbi").exists()
