"""ORACLE (test infrastructure): pure-Python restatement of the reference's evaluation stage.

  * extract_tp_fp_snp  -- program/extract_TP_FP_SNPs.py:12-57 (awk SNP/QUAL filter, `fgrep -wf` TP/FP split)
  * performance_row    -- scripts/caller_performance_compare.R:29-55,77-136 (TP/FP/FN, precision/recall/F1)

PARITY: PINNED.  tests/golden/eval/ holds outputs of the reference's own script run in the build container
(tests/golden/make_eval_golden.py); tests/test_eval_oracle.py checks this restatement against them byte for byte.
The R table has no runnable reference here (no R): its restatement follows the in-tree source; R's round()
is IEEE round-half-even on the scaled value (R 3.5.1, config/conda_env.yaml:7), reproduced with Python's round()."""
import os
import re

_WORD = re.compile(r"[A-Za-z0-9_]")
_NUM = re.compile(r"^[ \t]*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)[ \t]*$")


def _awk_ge_20_or_dot(f6):
    """awk: ($6>=20||$6=="."): numeric comparison when $6 looks like a number, string comparison otherwise"""
    if _NUM.match(f6):
        return float(f6) >= 20
    return f6 >= "20" or f6 == "."


def is_snp_line(fields, need_qual=True):
    if len(fields) < 5:
        return False
    if not (re.fullmatch(r"[ACGT]", fields[3]) and re.fullmatch(r"[ACGT]", fields[4])):
        return False
    if need_qual:
        f6 = fields[5] if len(fields) > 5 else ""
        return _awk_ge_20_or_dot(f6)
    return True


def truth_patterns(truth_lines):
    """awk '$4~/^[ACGT]$/&&$5~/^[ACGT]$/{print $2,".",$4,$5}' (header lines never pass: their $4 is not a base)"""
    pats = set()
    for ln in truth_lines:
        f = ln.rstrip("\n").split("\t")
        if is_snp_line(f, need_qual=False):
            pats.add((f[1], f[3], f[4]))
    return pats


def line_matches(line, pats):
    """`fgrep -w` of any "POS\\t.\\tREF\\tALT": the match must start after a non-word character (or at the line
    start) and end before one (or at the line end).  Patterns are digits TAB . TAB base TAB base, so a match
    covers the tail of one field, two whole fields and the head of a fourth."""
    f = line.split("\t")
    for i in range(len(f) - 3):
        if f[i + 1] != "." or len(f[i + 2]) != 1:
            continue
        m = re.search(r"(\d+)$", f[i])
        if not m:
            continue
        digits, start = m.group(1), m.start(1)
        # only the whole trailing digit run can start at a word boundary (a shorter suffix follows a digit)
        if start > 0 and _WORD.match(f[i][start - 1]):
            continue
        tail = f[i + 3]
        if not tail or (len(tail) > 1 and _WORD.match(tail[1])):
            continue
        if (digits, f[i + 2], tail[0]) in pats:
            return True
    return False


def extract_tp_fp_snp(vcf_file, snp_file):
    """same outputs, same paths as the reference function (fp/ must exist, as Snakemake makes it)"""
    dirname = os.path.dirname(vcf_file)
    fname_wo_ext = os.path.basename(vcf_file)[:-4]
    filtered_out = vcf_file[:-4] + ".filtered.vcf"
    fp_out = os.path.join(dirname, "fp", fname_wo_ext + ".fp.vcf")
    lines = open(vcf_file).read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    header = [ln for ln in lines if ln.startswith("#")]
    body = [ln for ln in lines if is_snp_line(ln.split("\t"))]
    with open(filtered_out, "w") as fh:
        fh.write("".join(x + "\n" for x in header + body))
    if os.path.basename(vcf_file).split(".")[0].endswith(("-1-0", "-0-1")):
        with open(fp_out, "w") as fh:
            fh.write("".join(x + "\n" for x in header + body))
        return
    os.makedirs(os.path.join(dirname, "tp"), exist_ok=True)
    tp_out = os.path.join(dirname, "tp", fname_wo_ext + ".tp.vcf")
    pats = truth_patterns(open(snp_file).read().split("\n"))
    hit = [line_matches(ln, pats) for ln in body]
    with open(tp_out, "w") as fh:
        fh.write("".join(x + "\n" for x in header + [b for b, h in zip(body, hit) if h]))
    with open(fp_out, "w") as fh:
        fh.write("".join(x + "\n" for x in header + [b for b, h in zip(body, hit) if not h]))


# ---- "bring your own data" variant (program/extract_TP_FP_SNPs.py:60-105; eval_variant_custom.smk) ----
def custom_truth_patterns(rows):
    """awk -F"\t" '$2!="."&&$3!="."{print $1,".",$2,$3}' on `show-snps -CTHIlr` rows (P1, ref base, query base, ...)"""
    pats = set()
    for ln in rows:
        f = ln.rstrip("\n").split("\t")
        f += [""] * (3 - len(f))
        if f[1] != "." and f[2] != ".":
            pats.add((f[0], f[1], f[2]))
    return pats


def line_matches_any(line, pats):
    """`fgrep -w` of "P1\t.\tX\tY" for arbitrary field texts: some occurrence of the fixed string starts at the line start or
    behind a non-word character and ends at the line end or before one"""
    for p0, x, y in pats:
        s = f"{p0}\t.\t{x}\t{y}"
        if not s:
            continue
        at = line.find(s)
        while at >= 0:
            before_ok = at == 0 or not _WORD.match(line[at - 1])
            end = at + len(s)
            after_ok = end == len(line) or not _WORD.match(line[end])
            if before_ok and after_ok:
                return True
            at = line.find(s, at + 1)
    return False


def extract_tp_fp_custom_snp(vcf_file, snp_file, outdir, caller):
    """same outputs as the reference function: <outdir>/<caller>.filtered.vcf, fp/<caller>.fp.vcf, tp/<caller>.tp.vcf
    (fp/ must exist)"""
    filtered_out = os.path.join(outdir, caller + ".filtered.vcf")
    fp_out = os.path.join(outdir, "fp", caller + ".fp.vcf")
    lines = open(vcf_file).read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    header = [ln for ln in lines if ln.startswith("#")]
    body = [ln for ln in lines if is_snp_line(ln.split("\t"))]

    def dump(path, rows):
        with open(path, "w") as fh:
            fh.write("".join(x + "\n" for x in header + rows))

    dump(filtered_out, body)
    if os.path.basename(vcf_file).split(".")[0].endswith(("-1-0", "-0-1")):
        dump(fp_out, body)
        return
    os.makedirs(os.path.join(outdir, "tp"), exist_ok=True)
    rows = open(snp_file).read().split("\n")
    if rows and rows[-1] == "":
        rows.pop()
    pats = custom_truth_patterns(rows)
    hit = [line_matches_any(ln, pats) for ln in body]
    dump(os.path.join(outdir, "tp", caller + ".tp.vcf"), [b for b, h in zip(body, hit) if h])
    dump(fp_out, [b for b, h in zip(body, hit) if not h])


CUSTOM_HEADER = ["caller", "genomediff", "calleridentify", "TP", "FP", "precision", "recall", "f1"]


def custom_performance_row(filtered_vcf, snp_file, caller):
    """one row of scripts/custom_snp_benchmark.R:23-27,41-88: truth = "P1-ref-alt" of the show-snps rows whose bases are
    not "." (comment lines '#' skipped, duplicates kept in the count), calls = rows with single A/C/G/T REF and ALT"""
    truth = []
    for ln in open(snp_file):
        if ln.startswith("#") or not ln.strip():
            continue
        f = ln.rstrip("\n").split("\t")
        if len(f) >= 3 and f[1] != "." and f[2] != ".":
            truth.append(f"{f[0]}-{f[1]}-{f[2]}")
    n_truth = len(truth)
    snp = make_snp_vector(filtered_vcf)
    n_rows = sum(1 for ln in open(filtered_vcf) if not ln.startswith("#") and ln.strip())
    if n_rows == 0:
        return [caller, str(n_truth), "0", "0", "0", "NA", "NA", "NA"]
    n_id = len(snp)
    s, t = set(snp), set(truth)
    tp, fp = len(s & t), len(s - t)

    def div(a, b):
        return float("nan") if b == 0 else a / b
    precision = r_round3(div(tp, n_id)) if n_id else float("nan")
    recall = r_round3(div(tp, n_truth)) if n_truth else float("nan")
    den = precision + recall
    f1 = r_round3(2 * (precision * recall) / den) if den == den and den != 0 else float("nan")
    fmt = lambda v: "NaN" if v != v else r_num(v)
    return [caller, str(n_truth), str(n_id), str(tp), str(fp), fmt(precision), fmt(recall), fmt(f1)]


# ---- scripts/caller_performance_compare.R ----
CALLER_MAP = {"bcftools": "BCFtools", "clc": "CLC", "freebayes": "FreeBayes", "gatk": "GATK", "lofreq": "LoFreq",
              "varscan": "VarScan2"}


def make_snp_vector(vcf):
    """paste(POS, REF, ALT, sep="-") of the rows whose REF and ALT are single A/C/G/T (R :29-55); duplicates kept"""
    out = []
    for ln in open(vcf):
        if ln.startswith("#") or not ln.strip():
            continue
        f = ln.rstrip("\n").split("\t")
        if len(f) >= 5 and f[3] in "ACGT" and len(f[3]) == 1 and f[4] in "ACGT" and len(f[4]) == 1:
            out.append(f"{f[1]}-{f[3]}-{f[4]}")
    return out


def r_round3(x):
    return round(x, 3)          # IEEE: decimal round-half-even on the binary value, as R >= 3.x sprintf-based round


def r_num(x):
    """R's write.table formatting of a double: up to 15 significant digits, no trailing zeros, NA"""
    if x is None:
        return "NA"
    if isinstance(x, int):
        return str(x)
    s = f"{x:.15g}"
    return s


def performance_row(filtered_vcf, truth_vectors, mix_samples):
    """one row of final_tables/caller_performance.tsv (R :77-136)"""
    parts = os.path.basename(filtered_vcf).split(".")
    sample, caller_lower = parts[0], parts[2]
    caller = CALLER_MAP.get(caller_lower, caller_lower)
    snp = make_snp_vector(filtered_vcf)
    n_id = len(snp)
    if sample in mix_samples:
        truth = truth_vectors[sample[:2]]
        n_truth = len(truth)
        if n_id > 0:
            s, t = set(snp), set(truth)
            tp, fp = len(s & t), len(s - t)
            precision = r_round3(tp / n_id)
            recall = r_round3(tp / n_truth) if n_truth else None
            if recall is None or precision + recall == 0:
                f1 = None               # R: 0/0 = NaN, written as NaN
                f1s = "NaN"
            else:
                f1 = r_round3(2 * (precision * recall) / (precision + recall))
                f1s = r_num(f1)
            return [caller, sample, str(n_truth), str(n_id), str(tp), str(fp), r_num(precision),
                    r_num(recall) if recall is not None else "NaN", f1s]
        return [caller, sample, str(n_truth), "0", "0", "0", "NA", "NA", "NA"]
    return [caller, sample, "0", str(n_id), "0", str(n_id), "0", "NA", "NA"]


TABLE_HEADER = ["caller", "mixture", "genomediff", "calleridentify", "TP", "FP", "Precision", "Recall", "F1"]
