"""Generates tests/golden/eval/*: inputs with every matcher quirk of SURVEY.md B.5 and the outputs of the
REFERENCE's own program/extract_TP_FP_SNPs.py run on them (needs /root/reference, bash, awk, fgrep: run in
the build container, outputs are committed).  Usage: python tests/golden/make_eval_golden.py"""
import os
import random
import shutil
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "eval")
REF_SCRIPT = "/root/reference/program/extract_TP_FP_SNPs.py"
HDR = ["##fileformat=VCFv4.2", "##source=golden", "##contig=<ID=Merlin_1555,length=245038>",
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"]


def truth_vcf(rng):
    rows = []
    pos = 50
    for _ in range(400):
        pos += rng.randint(1, 400)
        ref = rng.choice("ACGT")
        alt = rng.choice([b for b in "ACGT" if b != ref])
        rows.append(f"Merlin_1555\t{pos}\t.\t{ref}\t{alt}\t30\tPASS\tDP=30;TYPE=SNV")
    rows.append("Merlin_1555\t1100\t.\tA\tC\t30\tPASS\tDP=30;TYPE=SNV")
    rows.append("Merlin_1555\t5000\t.\tAT\tA\t30\tPASS\tDP=30;TYPE=INDEL")         # not a SNP pattern
    rows.append("Merlin_1555\t6000\t.\tA\tC,G\t30\tPASS\tDP=30;TYPE=SNV")          # multi-ALT: not a pattern
    rows.append("Merlin_1555\t7000\t.\tG\tT\t30\tPASS\tDP=30;TYPE=SNV")
    rows.append("Merlin_1555\t7000\t.\tG\tA\t30\tPASS\tDP=30;TYPE=SNV")
    return ["##fileformat=VCFv4.2", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"] + rows


def caller_vcf(rng, truth_rows):
    body = []
    snps = [r.split("\t") for r in truth_rows if not r.startswith("#")]
    for f in snps:
        u = rng.random()
        if u < 0.55 and len(f[3]) == 1 and len(f[4]) == 1:
            q = rng.choice(["30", "20", "19.9", "225.007", ".", "1e2", "5"])
            body.append(f"Merlin_1555\t{f[1]}\t.\t{f[3]}\t{f[4]}\t{q}\tPASS\tDP=88;AF=0.123;X\tGT\t1")
    for _ in range(150):                                                          # false positives
        pos = rng.randint(1, 245000)
        ref = rng.choice("ACGT")
        alt = rng.choice([b for b in "ACGT" if b != ref])
        body.append(f"Merlin_1555\t{pos}\t.\t{ref}\t{alt}\t{rng.choice(['33', '.', '12', '20.0'])}\tPASS\tDP=40\tGT\t1")
    # quirks
    body.append("Merlin_1555\t11100\t.\tA\tC\t50\tPASS\tDP=40\tGT\t1")      # POS 1100 is a suffix: no word match
    body.append("Merlin_1555\t1100\trs1\tA\tC\t50\tPASS\tDP=40\tGT\t1")     # ID is not "."
    body.append("Merlin_1555\t1100\t.\tA\tC\t50\tPASS\tDP=40\tGT\t1")       # the real hit
    body.append("Merlin_1555\t1100\t.\tA\tC\t60\tPASS\tDP=41\tGT\t1")       # duplicate call
    body.append("OtherChrom\t7000\t.\tG\tT\t50\tPASS\tDP=40\tGT\t1")        # CHROM is not compared
    body.append("Merlin_1555\t7000\t.\tG\tA\t50\tPASS\tDP=40\tGT\t1")
    body.append("Merlin_1555\t7000\t.\tG\tC\t50\tPASS\tDP=40\tGT\t1")       # same POS, other ALT: FP
    body.append("Merlin_1555\t6000\t.\tA\tC\t50\tPASS\tDP=40\tGT\t1")       # truth is multi-ALT: FP
    body.append("Merlin_1555\t5000\t.\tAT\tA\t50\tPASS\tDP=40\tGT\t1")      # indel: filtered out
    body.append("Merlin_1555\t8000\t.\ta\tC\t50\tPASS\tDP=40\tGT\t1")       # lowercase REF: filtered out
    body.append("Merlin_1555\t8001\t.\tA\tN\t50\tPASS\tDP=40\tGT\t1")       # N: filtered out
    body.append("Merlin_1555\t8002\t.\tA\tC,T\t50\tPASS\tDP=40\tGT\t1")     # multi-ALT: filtered out
    rng.shuffle(body)
    return HDR + body


def main():
    if not os.path.exists(REF_SCRIPT):
        sys.exit("needs /root/reference")
    rng = random.Random(20261018)
    shutil.rmtree(OUT, ignore_errors=True)
    truth = truth_vcf(rng)
    for sample in ("TM-1-1", "TA-1-0"):         # a mixture and a pure strain
        d = os.path.join(OUT, sample)
        os.makedirs(os.path.join(d, "fp"))
        vcf = os.path.join(d, f"{sample}.Merlin.bcftools.vcf")
        open(vcf, "w").write("\n".join(caller_vcf(rng, truth)) + "\n")
        tpath = os.path.join(d, "truth.vcf")
        open(tpath, "w").write("\n".join(truth) + "\n")
        subprocess.check_call([sys.executable, REF_SCRIPT, vcf, tpath, "hcmv", d, "bcftools"])
        time.sleep(1.0)                          # the script does not wait for its `tp` child
    # "bring your own data" mode: truth = show-snps -CTHIlr rows (P1, ref base, query base, P2, BUFF, DIST, R, Q, LEN R, LEN Q,
    # FRM, FRM, TAG R, TAG Q); "." marks an indel side
    d = os.path.join(OUT, "custom")
    os.makedirs(os.path.join(d, "fp"))
    rows = []
    for r in truth:
        if r.startswith("#"):
            continue
        f = r.split("\t")
        if len(f[3]) == 1 and len(f[4]) == 1:
            rows.append(f"{f[1]}\t{f[3]}\t{f[4]}\t{int(f[1]) + 17}\t12\t{f[1]}\t0\t0\t245038\t237683\t1\t1\tRefGenome\tQryGenome")
    rows.append("9000\t.\tA\t9017\t0\t9000\t0\t0\t245038\t237683\t1\t1\tRefGenome\tQryGenome")      # insertion: no pattern
    rows.append("9100\tC\t.\t9117\t0\t9100\t0\t0\t245038\t237683\t1\t1\tRefGenome\tQryGenome")      # deletion: no pattern
    rows.append("9200\tN\tA\t9217\t0\t9200\t0\t0\t245038\t237683\t1\t1\tRefGenome\tQryGenome")      # pattern no SNP line can match
    snp_rows = os.path.join(d, "genome_diff.snps")
    open(snp_rows, "w").write("\n".join(rows) + "\n")
    vcf = os.path.join(d, "mysample.calls.vcf")
    body = caller_vcf(rng, truth)
    body.append("Merlin_1555\t9200\t.\tA\tC\t50\tPASS\tDP=40\tGT\t1")
    open(vcf, "w").write("\n".join(body) + "\n")
    subprocess.check_call([sys.executable, REF_SCRIPT, vcf, snp_rows, "custom", d, "mycaller"])
    time.sleep(1.0)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
