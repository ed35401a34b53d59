"""Static check of the drop-in boundary: every `output:` / `log:` path the reference's rules `bwa`, `rmdup`,
`mpileup` and `bcftools` declare is one quasimodo_b200.rules knows and the driver command writes.  Parses the reference's own
rule files, so it only runs where /root/reference exists (the build container)."""
import os
import re

import pytest

from quasimodo_b200 import rules

REF = "/root/reference/rules"


def parse_rule(path, rule):
    """-> (outputs {name: (dir variable, pattern)}, log (dir variable, pattern) or None) of `rule` in a .smk file"""
    text = open(path).read()
    m = re.search(rf"^rule {rule}:\n(.*?)(?=^rule |\Z)", text, re.S | re.M)
    assert m, f"rule {rule} not found in {path}"
    body = m.group(1)

    def section(key):
        s = re.search(rf"^    {key}:\s*\n((?:^        .*\n|^\s*\n)*)", body, re.M)
        return s.group(1) if s else ""

    def entries(sec):
        out = {}
        for ln in sec.splitlines():
            ln = ln.split("#")[0].strip().rstrip(",")
            if not ln:
                continue
            e = re.match(r'(?:(\w+)\s*=\s*)?(\w+)\s*\+\s*"([^"]+)"$', ln)
            assert e, f"unparsed line in rule {rule}: {ln!r}"
            out[e.group(1) or ""] = (e.group(2), e.group(3))
        return out

    outs = entries(section("output"))
    log = entries(section("log"))
    bench = entries(section("benchmark"))
    return outs, (log.get("") if log else None), (bench.get("") if bench else None)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine")
def test_declared_outputs_of_the_replaced_rules_are_all_produced():
    where = {"bwa": "bwa.smk", "rmdup": "rmdup.smk", "mpileup": "vcfcall.smk", "bcftools": "vcfcall.smk"}
    for rule, f in where.items():
        outs, log, bench = parse_rule(os.path.join(REF, f), rule)
        assert outs == rules.RULE_OUTPUTS[rule], (rule, outs)
        assert log == rules.RULE_LOGS.get(rule), (rule, log)
        assert bench == rules.RULE_BENCHMARKS.get(rule), (rule, bench)
        for name in outs:
            assert (rule, name) in rules.DRIVER_OPTION, f"output {name!r} of rule {rule} has no driver option"
    # the shell lines' side files: `samtools index` after bwa and rmdup, `tabix -p vcf` after bcftools
    assert "samtools index {output.sortedbam}" in open(os.path.join(REF, "bwa.smk")).read()
    assert "samtools index {output.rmdupbam}" in open(os.path.join(REF, "rmdup.smk")).read()
    assert "tabix -p vcf {output.vcf_bgz}" in open(os.path.join(REF, "vcfcall.smk")).read()


def test_sample_command_names_every_path(tmp_path):
    dirs = {k: str(tmp_path / k) for k in ("seq_dir", "snpcall_dir", "report_dir")}
    argv, paths = rules.sample_command("qm_driver", dirs, "TM-1-1", "Merlin", "ref.fa", "a.fq", "b.fq")
    assert len(paths) == 6 + 3 and len(set(paths)) == len(paths)
    for opt in ("--bam", "--rmdup-bam", "--mpileup", "--vcf", "--metrics"):
        assert argv[argv.index(opt) + 1] in paths
    assert str(tmp_path / "seq_dir" / "bam" / "TM-1-1.Merlin.bam") in paths
    assert str(tmp_path / "snpcall_dir" / "bcftools" / "TM-1-1.Merlin.bcftools.vcf.gz.tbi") in paths
    argv_b, paths_b = rules.sample_command("qm_driver", dirs, "TM-1-1", "Merlin", "ref.fa", "a.fq", "b.fq", benchmark=True)
    assert paths_b == paths and argv_b[:len(argv)] == argv
    assert argv_b[-2:] == ["--benchmark", str(tmp_path / "report_dir" / "benchmarks" / "TM-1-1.Merlin.bcftools.benchmark.txt")]
