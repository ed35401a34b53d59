"""CPU: the evaluation oracle (oracle/eval_py.py) against the golden outputs of the reference's own
program/extract_TP_FP_SNPs.py (tests/golden/eval, made by tests/golden/make_eval_golden.py), byte for byte;
plus the caller_performance table restatement on the same files."""
import filecmp
import os
import shutil

import pytest

from oracle import eval_py

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval")


@pytest.mark.parametrize("sample", ["TM-1-1", "TA-1-0"])
def test_extract_matches_reference_script(tmp_path, sample):
    src = os.path.join(GOLD, sample)
    d = tmp_path / sample
    os.makedirs(d / "fp")
    name = f"{sample}.Merlin.bcftools"
    shutil.copy(os.path.join(src, name + ".vcf"), d)
    eval_py.extract_tp_fp_snp(str(d / (name + ".vcf")), os.path.join(src, "truth.vcf"))
    assert filecmp.cmp(d / (name + ".filtered.vcf"), os.path.join(src, name + ".filtered.vcf"), shallow=False)
    assert filecmp.cmp(d / "fp" / (name + ".fp.vcf"), os.path.join(src, "fp", name + ".fp.vcf"), shallow=False)
    if sample == "TM-1-1":
        assert filecmp.cmp(d / "tp" / (name + ".tp.vcf"), os.path.join(src, "tp", name + ".tp.vcf"), shallow=False)
    else:
        assert not (d / "tp").exists()


def test_custom_mode_matches_reference_script(tmp_path):
    """the "bring your own data" variant (truth = show-snps rows; program/extract_TP_FP_SNPs.py:60-105) against the files the
    reference's script wrote in that mode"""
    src = os.path.join(GOLD, "custom")
    d = tmp_path / "custom"
    os.makedirs(d / "fp")
    eval_py.extract_tp_fp_custom_snp(os.path.join(src, "mysample.calls.vcf"), os.path.join(src, "genome_diff.snps"), str(d), "mycaller")
    for rel in ("mycaller.filtered.vcf", os.path.join("fp", "mycaller.fp.vcf"), os.path.join("tp", "mycaller.tp.vcf")):
        assert filecmp.cmp(d / rel, os.path.join(src, rel), shallow=False), rel
    tp = open(os.path.join(src, "tp", "mycaller.tp.vcf")).read()
    fp = open(os.path.join(src, "fp", "mycaller.fp.vcf")).read()
    assert "\t9200\t.\tA\tC\t" in fp and "\t1100\t.\tA\tC\t" in tp          # an N row of the truth matches nothing
    row = eval_py.custom_performance_row(os.path.join(src, "mycaller.filtered.vcf"), os.path.join(src, "genome_diff.snps"), "mycaller")
    assert row[0] == "mycaller" and int(row[3]) + int(row[4]) <= int(row[2]) and 0 < float(row[5]) < 1


def test_quirks_present_in_golden():
    """the golden TP file holds the hits the script's word matching implies"""
    tp = open(os.path.join(GOLD, "TM-1-1", "tp", "TM-1-1.Merlin.bcftools.tp.vcf")).read()
    fp = open(os.path.join(GOLD, "TM-1-1", "fp", "TM-1-1.Merlin.bcftools.fp.vcf")).read()
    assert "\t11100\t.\tA\tC\t" in fp and "\t11100\t" not in tp          # 1100 is not a word inside 11100
    assert "\t1100\trs1\tA\tC\t" in fp                                    # ID must be "."
    assert tp.count("\t1100\t.\tA\tC\t") >= 2                            # duplicates both kept
    assert "OtherChrom\t7000\t.\tG\tT\t" in tp                            # CHROM is not compared
    assert "\t7000\t.\tG\tC\t" in fp and "\t6000\t.\tA\tC\t" in fp


def test_performance_row():
    src = os.path.join(GOLD, "TM-1-1")
    truth = {"TM": eval_py.make_snp_vector(os.path.join(src, "truth.vcf"))}
    row = eval_py.performance_row(os.path.join(src, "TM-1-1.Merlin.bcftools.filtered.vcf"), truth, ["TM-1-1"])
    snp = eval_py.make_snp_vector(os.path.join(src, "TM-1-1.Merlin.bcftools.filtered.vcf"))
    tp = len(set(snp) & set(truth["TM"]))
    assert row[:2] == ["BCFtools", "TM-1-1"] and int(row[3]) == len(snp) and int(row[4]) == tp
    assert int(row[5]) == len(set(snp) - set(truth["TM"]))
    assert row[6] == eval_py.r_num(round(tp / len(snp), 3))
    pure = eval_py.performance_row(os.path.join(GOLD, "TA-1-0", "TA-1-0.Merlin.bcftools.filtered.vcf"), truth, ["TM-1-1"])
    assert pure[2] == "0" and pure[4] == "0" and pure[5] == pure[3] and pure[6:] == ["0", "NA", "NA"]


def test_python_round_equals_r_long_double_rounding_on_every_ratio():
    """R 3.5.1's round(x, 3) is nearbyintl(x * 10^3) / 10^3 in long double (round half even on the SCALED value), the
    product's and the oracle's tables use Python's round(x, 3) (decimal round-half-even on the BINARY value).  The two can only differ
    when the scaled value lands on a half within a long-double ulp; on every precision / recall a table can hold (TP / n with
    n <= 400 here; all n < 1200 were checked when this test was written) and on every F1 built from such pairs they agree."""
    import numpy as np
    if np.finfo(np.longdouble).nmant < 63:
        import pytest
        pytest.skip("no 80-bit long double on this host")
    k = np.longdouble(1000)
    vals = []
    for b in range(1, 401):
        vals.append(np.arange(0, b + 1) / b)
    x = np.concatenate(vals)
    r = (np.rint(x.astype(np.longdouble) * k) / k).astype(np.float64)
    p = np.array([round(float(v), 3) for v in x])
    assert np.array_equal(r, p)
    # F1 = 2 p r / (p + r) of rounded p, r (scripts/caller_performance_compare.R:51)
    g = np.round(np.linspace(0.001, 1, 250), 3)
    pp, rr = np.meshgrid(g, g)
    f = (2 * pp * rr / (pp + rr)).ravel()
    rf = (np.rint(f.astype(np.longdouble) * k) / k).astype(np.float64)
    pf = np.array([round(float(v), 3) for v in f])
    assert np.array_equal(rf, pf)
