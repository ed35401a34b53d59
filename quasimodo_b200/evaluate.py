"""Evaluation stage, host side: the mirror of the reference's program/extract_TP_FP_SNPs.py (same function
names, arguments, output files) and of the table part of scripts/caller_performance_compare.R, with the
matching itself done by the CUDA matcher (qm_eval_match) on packed (POS, REF, ALT) keys.

Text handling (reading VCF lines, the awk-equivalent SNP/QUAL filter, writing the surviving original lines
in original order) is host work; which line is TP / FP and which truth SNP is FN is decided on the device.
There is no CPU fallback: without the CUDA library / a B200 these functions raise."""
import ctypes as C
import os
import re

import numpy as np

from . import _lib
from .api import _check

_BASE = {"A": 0, "C": 1, "G": 2, "T": 3}
_NUM = re.compile(r"^[ \t]*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)[ \t]*$")
NO_KEY = np.uint64(0xFFFFFFFFFFFFFFFF)          # never equals a truth key


def _qual_ok(f6):
    """awk ($6>=20||$6=="."), program/extract_TP_FP_SNPs.py:24"""
    if _NUM.match(f6):
        return float(f6) >= 20
    return f6 >= "20" or f6 == "."


def _is_snp(f):
    return len(f) >= 5 and f[3] in _BASE and f[4] in _BASE


def snp_key(pos, ref, alt):
    return np.uint64((int(pos) << 8) | (_BASE[ref] << 4) | _BASE[alt])


def call_keys(body_lines):
    """key of each filtered caller line; lines the script's pattern can never match (ID not ".", POS not a
    plain number) get NO_KEY"""
    keys = np.empty(len(body_lines), dtype=np.uint64)
    for i, ln in enumerate(body_lines):
        f = ln.split("\t")
        if f[2] == "." and f[1].isdigit() and f[1].isascii():
            keys[i] = snp_key(f[1], f[3], f[4])
        else:
            keys[i] = NO_KEY
    return keys


def truth_keys(snp_file):
    """keys of the truth VCF rows with single-base REF and ALT (the script's awk pattern generator, :47)"""
    keys = []
    for ln in open(snp_file):
        f = ln.rstrip("\n").split("\t")
        if _is_snp(f) and f[1].isdigit():
            keys.append(snp_key(f[1], f[3], f[4]))
    return np.array(keys, dtype=np.uint64)


def match_keys(ctx, ckeys, tkeys):
    """-> (call_flags, truth_flags) uint8 arrays, computed by the CUDA matcher"""
    import torch
    dev = torch.device(f"cuda:{ctx.device}")
    d_c = torch.from_numpy(ckeys.view(np.int64)).to(dev)
    d_t = torch.from_numpy(tkeys.view(np.int64)).to(dev)
    d_cf = torch.zeros(max(len(ckeys), 1), dtype=torch.uint8, device=dev)
    d_tf = torch.zeros(max(len(tkeys), 1), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = _lib.lib().qm_eval_match(ctx._h, C.c_void_p(d_c.data_ptr()), len(ckeys), C.c_void_p(d_t.data_ptr()), len(tkeys),
                                  C.c_void_p(d_cf.data_ptr()), C.c_void_p(d_tf.data_ptr()), C.c_void_p(st))
    _check(ctx._h, rc, "qm_eval_match")
    torch.cuda.synchronize(dev)
    return d_cf.cpu().numpy()[:len(ckeys)], d_tf.cpu().numpy()[:len(tkeys)]


def extract_tp_fp_snp(ctx, vcf_file, snp_file):
    """program/extract_TP_FP_SNPs.py:12-57 -- writes <vcf>.filtered.vcf, fp/<name>.fp.vcf and (mixtures)
    tp/<name>.tp.vcf next to vcf_file; returns (n_filtered, n_tp, n_fp)"""
    dirname = os.path.dirname(vcf_file)
    fname_wo_ext = os.path.basename(vcf_file)[:-4]
    filtered_out = vcf_file[:-4] + ".filtered.vcf"
    os.makedirs(os.path.join(dirname, "fp"), exist_ok=True)
    fp_out = os.path.join(dirname, "fp", fname_wo_ext + ".fp.vcf")
    lines = open(vcf_file).read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    header = [ln for ln in lines if ln.startswith("#")]
    body = []
    for ln in lines:
        f = ln.split("\t")
        if _is_snp(f) and _qual_ok(f[5] if len(f) > 5 else ""):
            body.append(ln)

    def dump(path, rows):
        with open(path, "w") as fh:
            fh.write("".join(x + "\n" for x in header + rows))

    dump(filtered_out, body)
    if os.path.basename(vcf_file).split(".")[0].endswith(("-1-0", "-0-1")):
        dump(fp_out, body)
        return len(body), 0, len(body)
    os.makedirs(os.path.join(dirname, "tp"), exist_ok=True)
    tp_out = os.path.join(dirname, "tp", fname_wo_ext + ".tp.vcf")
    flags, _ = match_keys(ctx, call_keys(body), truth_keys(snp_file))
    tp = [b for b, h in zip(body, flags) if h]
    fp = [b for b, h in zip(body, flags) if not h]
    dump(tp_out, tp)
    dump(fp_out, fp)
    return len(body), len(tp), len(fp)


# ---- "bring your own data" variant: truth = `show-snps -CTHIlr` rows (program/extract_TP_FP_SNPs.py:60-105, rule
# eval_variant_custom.smk:58-92; table scripts/custom_snp_benchmark.R:23-27,41-88) ----
def custom_truth_keys(snp_file):
    """keys of the rows the script's pattern generator ($2!="."&&$3!=".") emits AND a filtered caller line can match:
    a caller line's REF / ALT are single A/C/G/T and its POS is a plain number, so only such patterns can ever hit"""
    keys = []
    for ln in open(snp_file):
        f = ln.rstrip("\n").split("\t")
        if len(f) >= 3 and f[1] in _BASE and f[2] in _BASE and f[0].isdigit() and f[0].isascii() and (f[0] == "0" or f[0][0] != "0"):
            keys.append(snp_key(f[0], f[1], f[2]))
    return np.array(keys, dtype=np.uint64)


def extract_tp_fp_custom_snp(ctx, vcf_file, snp_file, outdir, caller):
    """program/extract_TP_FP_SNPs.py:60-105 -- writes <outdir>/<caller>.filtered.vcf, fp/<caller>.fp.vcf and (mixtures)
    tp/<caller>.tp.vcf; returns (n_filtered, n_tp, n_fp)"""
    filtered_out = os.path.join(outdir, caller + ".filtered.vcf")
    os.makedirs(os.path.join(outdir, "fp"), exist_ok=True)
    fp_out = os.path.join(outdir, "fp", caller + ".fp.vcf")
    lines = open(vcf_file).read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    header = [ln for ln in lines if ln.startswith("#")]
    body = []
    for ln in lines:
        f = ln.split("\t")
        if _is_snp(f) and _qual_ok(f[5] if len(f) > 5 else ""):
            body.append(ln)

    def dump(path, rows):
        with open(path, "w") as fh:
            fh.write("".join(x + "\n" for x in header + rows))

    dump(filtered_out, body)
    if os.path.basename(vcf_file).split(".")[0].endswith(("-1-0", "-0-1")):
        dump(fp_out, body)
        return len(body), 0, len(body)
    os.makedirs(os.path.join(outdir, "tp"), exist_ok=True)
    flags, _ = match_keys(ctx, call_keys(body), custom_truth_keys(snp_file))
    tp = [b for b, h in zip(body, flags) if h]
    fp = [b for b, h in zip(body, flags) if not h]
    dump(os.path.join(outdir, "tp", caller + ".tp.vcf"), tp)
    dump(fp_out, fp)
    return len(body), len(tp), len(fp)


CUSTOM_TABLE_HEADER = ["caller", "genomediff", "calleridentify", "TP", "FP", "precision", "recall", "f1"]


def custom_performance_row(ctx, filtered_vcf, snp_file, caller):
    """one row of the table of scripts/custom_snp_benchmark.R; set sizes from the device matcher.  A truth row whose bases
    are not single A/C/G/T can never equal a call: it only counts in genomediff."""
    truth, n_truth = [], 0
    for ln in open(snp_file):
        if ln.startswith("#") or not ln.strip():
            continue
        f = ln.rstrip("\n").split("\t")
        if len(f) >= 3 and f[1] != "." and f[2] != ".":
            n_truth += 1
            if f[1] in _BASE and f[2] in _BASE:
                truth.append((f[0], f[1], f[2]))
    snp = _snp_rows(filtered_vcf)
    n_rows = sum(1 for ln in open(filtered_vcf) if not ln.startswith("#") and ln.strip())
    if n_rows == 0:
        return [caller, str(n_truth), "0", "0", "0", "NA", "NA", "NA"]
    n_id = len(snp)
    uniq_c, uniq_t = sorted(set(snp)), sorted(set(truth))
    # R compares the strings "POS-REF-ALT": a position text that is not a plain number only ever equals itself
    num = lambda p: p.isdigit() and p.isascii() and (p == "0" or p[0] != "0")
    plain_c, plain_t = [x for x in uniq_c if num(x[0])], [x for x in uniq_t if num(x[0])]
    odd = len(set(uniq_c) - set(plain_c) & set(uniq_t) - set(plain_t)) if len(plain_c) != len(uniq_c) else 0
    ck = np.array([snp_key(int(p), r, a) for p, r, a in plain_c], dtype=np.uint64)
    tk = np.array([snp_key(int(p), r, a) for p, r, a in plain_t], dtype=np.uint64)
    cf, _ = match_keys(ctx, ck, tk)
    tp = int(cf.sum()) + odd
    fp = len(uniq_c) - tp
    nan = float("nan")
    precision = round(tp / n_id, 3) if n_id else nan
    recall = round(tp / n_truth, 3) if n_truth else nan
    den = precision + recall
    f1 = round(2 * (precision * recall) / den, 3) if den == den and den != 0 else nan
    return [caller, str(n_truth), str(n_id), str(tp), str(fp), _r_num(precision), _r_num(recall), _r_num(f1)]


def write_custom_benchmark_table(path, rows):
    with open(path, "w") as fh:
        fh.write("\t".join(CUSTOM_TABLE_HEADER) + "\n")
        for r in rows:
            fh.write("\t".join(r) + "\n")


# ---- table part of scripts/caller_performance_compare.R:29-55,77-143 ----
CALLER_MAP = {"bcftools": "BCFtools", "clc": "CLC", "freebayes": "FreeBayes", "gatk": "GATK", "lofreq": "LoFreq",
              "varscan": "VarScan2"}
TABLE_HEADER = ["caller", "mixture", "genomediff", "calleridentify", "TP", "FP", "Precision", "Recall", "F1"]


def _snp_rows(vcf):
    """(POS, REF, ALT) of the rows make_snp_vector keeps (R :29-55); duplicates kept"""
    out = []
    for ln in open(vcf):
        if ln.startswith("#") or not ln.strip():
            continue
        f = ln.rstrip("\n").split("\t")
        if _is_snp(f):
            out.append((f[1], f[3], f[4]))
    return out


def _r_num(x):
    if x is None:
        return "NA"
    if x != x:
        return "NaN"
    return f"{x:.15g}"


def performance_row(ctx, filtered_vcf, truth_vcf_by_mix, mix_samples):
    """one row of final_tables/caller_performance.tsv; set sizes (TP/FP/FN) come from the device matcher on the
    de-duplicated keys, exactly the intersect / setdiff of the R script.  Returns (row, fn_count)."""
    parts = os.path.basename(filtered_vcf).split(".")
    sample, caller_lower = parts[0], parts[2]
    caller = CALLER_MAP.get(caller_lower, caller_lower)
    snp = _snp_rows(filtered_vcf)
    n_id = len(snp)
    if sample not in mix_samples:
        return [caller, sample, "0", str(n_id), "0", str(n_id), "0", "NA", "NA"], 0
    truth = _snp_rows(truth_vcf_by_mix[sample[:2]])
    n_truth = len(truth)
    if n_id == 0:
        return [caller, sample, str(n_truth), "0", "0", "0", "NA", "NA", "NA"], 0
    # R compares the strings "POS-REF-ALT": distinct strings <=> distinct (POS text, REF, ALT)
    uniq_c = sorted(set(snp))
    uniq_t = sorted(set(truth))
    numeric = all(p.isdigit() for p, _, _ in uniq_c + uniq_t)
    if not numeric:
        raise ValueError("non-numeric POS in a SNP row")
    ck = np.array([snp_key(int(p), r, a) for p, r, a in uniq_c], dtype=np.uint64)
    tk = np.array([snp_key(int(p), r, a) for p, r, a in uniq_t], dtype=np.uint64)
    cf, tf = match_keys(ctx, ck, tk)
    tp, fp, fn = int(cf.sum()), int(len(ck) - cf.sum()), int(len(tk) - tf.sum())
    precision = round(tp / n_id, 3)
    recall = round(tp / n_truth, 3) if n_truth else float("nan")
    f1 = round(2 * (precision * recall) / (precision + recall), 3) if (precision + recall) == (precision + recall) and precision + recall != 0 else float("nan")
    return [caller, sample, str(n_truth), str(n_id), str(tp), str(fp), _r_num(precision), _r_num(recall), _r_num(f1)], fn


def write_performance_table(path, rows):
    with open(path, "w") as fh:
        fh.write("\t".join(TABLE_HEADER) + "\n")
        for r in rows:
            fh.write("\t".join(r) + "\n")
