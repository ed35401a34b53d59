"""Helpers shared by the driver tests: FASTA / FASTQ writers for simulated inputs, the driver binary."""
import gzip
import re
import subprocess

import numpy as np

from oracle.sort_py import reg2bin as _reg2bin
from quasimodo_b200 import build

BASES = "ACGTN"


def driver_path():
    build.build()
    return build.DRIVER


def run_driver(args, check=True):
    p = subprocess.run([driver_path()] + [str(a) for a in args], capture_output=True, text=True)
    if check and p.returncode != 0:
        raise AssertionError(f"qm_driver {' '.join(map(str, args))} -> {p.returncode}\n{p.stderr}")
    return p


def write_fasta(genome, path, width=70):
    with open(path, "w") as fh:
        off = 0
        for name, ln in zip(genome.names, genome.lens):
            fh.write(f">{name} test contig\n")
            s = "".join(BASES[c] for c in genome.codes[off:off + ln])
            for i in range(0, ln, width):
                fh.write(s[i:i + width] + "\n")
            off += ln


def pair_names(prefix, n):
    return [f"{prefix}.{i}" for i in range(n)]


def write_fastq(codes, quals, lens, names, path1, path2, gz=False, suffix=True, comment=""):
    op = (lambda p: gzip.open(p, "wt")) if gz else (lambda p: open(p, "w"))
    lut = np.frombuffer(BASES.encode(), dtype=np.uint8)
    with op(path1) as f1, op(path2) as f2:
        for i, nm in enumerate(names):
            for m, fh in ((0, f1), (1, f2)):
                r = 2 * i + m
                s = lut[codes[r, :lens[r]]].tobytes().decode()
                q = (quals[r, :lens[r]] + 33).astype(np.uint8).tobytes().decode()
                fh.write(f"@{nm}{'/%d' % (m + 1) if suffix else ''}{comment}\n{s}\n+\n{q}\n")


def revcomp_codes(c):
    c = c[::-1]
    return np.where(c < 4, 3 - c, 4).astype(np.uint8)


def check_bam_records(case):
    """every BAM record against the alignment record it was written from (case: bam, alns, perm, names, codes, quals, lens, W)"""
    bam, alns, perm = case["bam"], case["alns"], case["perm"]
    n_pairs = len(alns) // 2
    n_md = 0
    n_unmapped_rev = 0
    for rec, gi in zip(bam.records, perm):
        a = alns[gi]
        assert rec["name"] == case["names"][gi // 2]
        for f, g in (("rid", "rid"), ("pos", "pos"), ("flag", "flag"), ("mapq", "mapq"), ("mrid", "mate_rid"), ("mpos", "mate_pos"), ("tlen", "tlen")):
            assert rec[f] == int(a[g]), (gi, f)
        mapped = not (a["flag"] & 4)
        n_cig = int(a["n_cigar"]) if mapped else 0
        assert rec["cigar"] == [int(x) for x in a["cigar"][:n_cig]]
        L = case["lens"][gi]
        c, q = case["codes"][gi, :L], case["quals"][gi, :L]
        if a["flag"] & 0x10:       # flag alone, as every BAM -> FASTQ extractor does (an unmapped read placed at its reverse mate too)
            c, q = revcomp_codes(c), q[::-1]
            n_unmapped_rev += not mapped
        assert rec["seq"] == "".join("ACGTN"[x] for x in c)
        assert np.array_equal(rec["qual"], q)
        rlen = sum(x >> 4 for x in rec["cigar"] if (x & 15) in (0, 2))
        assert rec["bin"] == _reg2bin(rec["pos"], rec["pos"] + max(rlen, 1))
        assert rec["tags"].get("AS") == int(a["score"]) and rec["tags"].get("XS") == int(a["sub"])
        if n_cig:
            assert rec["tags"]["NM"] == int(a["nm"])
            # MD + CIGAR + SEQ must reproduce the reference, and NM = mismatches + inserted + deleted bases
            W = case["W"]
            off = int(np.concatenate([[0], np.cumsum(W.ref.lens)])[rec["rid"]]) + rec["pos"]
            md = re.findall(r"(\d+)|(\^[ACGTN]+)|([ACGTN])", rec["tags"]["MD"])
            toks = []
            for num, dele, mis in md:
                toks.append(("=", int(num)) if num else ("^", dele[1:]) if dele else ("x", mis))
            x, ref_out, ti, left, nm = 0, [], 0, 0, 0
            for cg in rec["cigar"]:
                op, ln = cg & 15, cg >> 4
                if op == 0:
                    k = 0
                    while k < ln:
                        if left == 0:
                            t = toks[ti]
                            ti += 1
                            if t[0] == "=":
                                left = t[1]
                                continue
                            assert t[0] == "x"
                            ref_out.append(t[1])
                            nm += 1
                            x += 1
                            k += 1
                            continue
                        step = min(left, ln - k)
                        ref_out.append(rec["seq"][x:x + step])
                        x += step
                        k += step
                        left -= step
                elif op == 2:
                    while left == 0 and toks[ti][0] == "=":
                        left = toks[ti][1]
                        ti += 1
                    assert left == 0 and toks[ti][0] == "^" and len(toks[ti][1]) == ln
                    ref_out.append(toks[ti][1])
                    ti += 1
                    nm += ln
                elif op == 1:
                    x += ln
                    nm += ln
                elif op == 4:
                    x += ln
            want = "".join("ACGT"[b] for b in W.ref.codes[off:off + rlen])
            assert "".join(ref_out) == want, (gi, rec["tags"]["MD"])
            assert nm == rec["tags"]["NM"]
            n_md += 1
        mate = alns[gi ^ 1]
        if not (mate["flag"] & 4) and 0 < mate["n_cigar"] < 255:
            assert rec["tags"]["MC"] == "".join(f"{int(x) >> 4}{'MIDNSHP=X'[int(x) & 15]}" for x in mate["cigar"][:mate["n_cigar"]])
        else:
            assert "MC" not in rec["tags"]
    return n_md


