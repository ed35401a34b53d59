#!/usr/bin/env python
"""File-level throughput (FASTQ -> BAM + BAI + count TSV + VCF + .vcf.gz + .tbi, pairs/s by the wall clock of
`qm_driver sample`, process start to exit): what a Snakemake job of the replaced rules would take.
usage: file_level_bench.py [pairs=1000000] [config2 sample index=4] [threads=16]
Writes plain FASTQ (fixed-width records assembled with numpy) and the reference FASTA into a temporary directory, runs the driver
twice (every output / counts + VCF only) and prints one JSON line per run."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quasimodo_b200 import build, workloads
from tests import drvutil

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
i = int(sys.argv[2]) if len(sys.argv) > 2 else 4
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 16
W = workloads.config2(i, n)
L = 150
t0 = time.perf_counter()
codes, quals, _, _ = W.simulate_host(0, n)
lut = np.frombuffer(b"ACGTN", dtype=np.uint8)


def write_fastq(path, rows, mate):
    m = len(rows)
    name = np.char.add(np.char.add("@sim.", np.char.zfill(np.arange(m).astype(str), 9)), f"/{mate}\n").astype("S")
    nw = name.dtype.itemsize
    rec = np.empty((m, nw + L + 3 + L + 1), dtype=np.uint8)
    rec[:, :nw] = np.frombuffer(name.tobytes(), dtype=np.uint8).reshape(m, nw)
    rec[:, nw:nw + L] = lut[codes[rows]]
    rec[:, nw + L:nw + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, nw + L + 3:nw + 2 * L + 3] = quals[rows] + 33
    rec[:, -1] = 10
    rec.tofile(path)


with tempfile.TemporaryDirectory() as d:
    fa, r1, r2 = os.path.join(d, "ref.fa"), os.path.join(d, "r1.fq"), os.path.join(d, "r2.fq")
    drvutil.write_fasta(W.ref, fa)
    write_fastq(r1, np.arange(0, 2 * n, 2), 1)
    write_fastq(r2, np.arange(1, 2 * n, 2), 2)
    fq_bytes = os.path.getsize(r1) + os.path.getsize(r2)
    print(f"inputs: {n} pairs, {fq_bytes / 1e6:.0f} MB of FASTQ, written in {time.perf_counter() - t0:.1f} s", file=sys.stderr)
    drv = build.DRIVER
    runs = [("bam+bai+counts+vcf+vcf.gz+tbi", ["--bam", os.path.join(d, "s.bam"), "--counts", os.path.join(d, "s.tsv"), "--vcf", os.path.join(d, "s.vcf")]),
            ("counts+vcf+vcf.gz+tbi", ["--counts", os.path.join(d, "t.tsv"), "--vcf", os.path.join(d, "t.vcf")]),
            ("vcf+vcf.gz+tbi", ["--vcf", os.path.join(d, "u.vcf")])]
    for what, outs in runs:
        t = time.perf_counter()
        p = subprocess.run([drv, "sample", "--ref", fa, "--r1", r1, "--r2", r2, "--sample", W.name.split(":")[1], "-t", str(threads)] + outs,
                           capture_output=True, text=True)
        dt = time.perf_counter() - t
        if p.returncode:
            sys.exit(f"qm_driver failed ({p.returncode}):\n{p.stderr}")
        sizes = {os.path.basename(o): os.path.getsize(o) for o in outs if not o.startswith("--")}
        line = [ln for ln in p.stderr.split("\n") if "from the first read to the last record" in ln]
        print(json.dumps({"metric": "file_level_read_pairs_per_s", "outputs": what, "pairs": n, "wall_s": round(dt, 2), "value": round(n / dt),
                          "unit": "pairs/s", "threads": threads, "fastq_mb": round(fq_bytes / 1e6), "output_bytes": sizes,
                          "driver_log": line[0].strip() if line else None,
                          "phases": next((ln.strip() for ln in p.stderr.split("\n") if "phases (s)" in ln), None), "workload": W.name}))
