"""Truth sets for the bundled strain pairs, made with the REFERENCE's own program/mummer2vcf.py.

nucmer / show-snps (mummer 3.23, rules/genome_diff.smk:20-22) are not in the image, so the `show-snps -CTHlr`
rows mummer2vcf.py consumes are derived from the bundled three-strain MAFFT alignment
(ref/msa/3_HCMV_ref.mafft.fas) by pairwise column projection (last row).  mummer2vcf.py then runs
unchanged (its only Biopython use, SeqIO.parse, is served by a 15-line shim on PYTHONPATH).
Outputs (committed): quasimodo_b200/data/truth/{TM,TA}.maskrepeat.variants.vcf.gz
Needs /root/reference: run in the build container.  Usage: python tests/golden/make_truth.py"""
import gzip
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(ROOT, "quasimodo_b200", "data", "truth")

SHIM = '''
class _Rec:
    def __init__(self, id, seq):
        self.id, self.seq = id, seq
def parse(path, fmt):
    name, chunks = None, []
    for ln in open(path):
        ln = ln.rstrip("\\n\\r")
        if ln.startswith(">"):
            if name is not None:
                yield _Rec(name, "".join(chunks))
            name, chunks = ln[1:].split()[0], []
        elif ln:
            chunks.append(ln)
    if name is not None:
        yield _Rec(name, "".join(chunks))
'''


def read_fasta(path):
    seqs, name = {}, None
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith(">"):
            name = ln[1:].split()[0]
            seqs[name] = []
        elif ln:
            seqs[name].append(ln)
    return {k: "".join(v) for k, v in seqs.items()}


WINDOW, MIN_IDENTITY = 200, 0.85


def alignable_columns(aref, aqry):
    """nucmer only reports differences inside the alignments it finds: maximal-match clusters extended while the identity holds.
    The MSA, being global, also pairs up the hypervariable loci (RL12-14, UL73/74, UL139-146 ...) column by column, where
    the strains are < 70 % identical and nucmer has no alignment at all -- projected as they are, those columns put 90 % of the
    truth SNPs into 67 one-kb windows with > 100 SNPs per kb.  A column counts as alignable when the WINDOW columns around it
    (gap-gap columns skipped) are at least MIN_IDENTITY identical."""
    cols = [(r, q) for r, q in zip(aref.upper(), aqry.upper()) if not (r == "-" and q == "-")]
    same = [1 if r == q else 0 for r, q in cols]
    pre = [0]
    for v in same:
        pre.append(pre[-1] + v)
    ok, h = [], WINDOW // 2
    for i in range(len(cols)):
        lo, hi = max(0, i - h), min(len(cols), i + h)
        ok.append((pre[hi] - pre[lo]) >= MIN_IDENTITY * (hi - lo))
    return ok


def rows_from_msa(aref, aqry, ref_name, qry_name, ref_len, qry_len):
    """show-snps -CTHlr style rows: P1 SUB SUB P2 BUFF DIST LENR LENQ FRM FRM TAGR TAGQ"""
    rows, p1, p2 = [], 0, 0
    ok = iter(alignable_columns(aref, aqry))
    for r, q in zip(aref.upper(), aqry.upper()):
        if r != "-":
            p1 += 1
        if q != "-":
            p2 += 1
        if r == "-" and q == "-":
            continue
        if not next(ok):
            continue                                   # inside a stretch nucmer would not align
        if r == q:
            continue
        if p1 == 0 or p2 == 0:
            continue                                   # unaligned leading overhang
        sub_r = "." if r == "-" else r
        sub_q = "." if q == "-" else q
        rows.append("\t".join([str(p1), sub_r, sub_q, str(p2), "0", "0", str(ref_len), str(qry_len), "1", "1", ref_name, qry_name]))
    return rows


def main():
    if not os.path.exists(REF):
        sys.exit("needs /root/reference")
    msa = read_fasta(os.path.join(REF, "ref", "msa", "3_HCMV_ref.mafft.fas"))
    refs = {"Merlin": "Merlin.BAC.fa", "TB40E": "TB40E.GFP.fa", "AD169": "AD169.BAC.fa"}
    genomes = {k: read_fasta(os.path.join(REF, "ref", v)) for k, v in refs.items()}
    # map MSA rows to bundled genomes by de-gapped identity
    row_of = {}
    for k, g in genomes.items():
        (gname, gseq), = g.items()
        for mname, mseq in msa.items():
            if mseq.replace("-", "").upper() == gseq.upper():
                row_of[k] = (mname, gname, len(gseq))
    assert len(row_of) == 3, f"MSA rows do not match the bundled genomes: {list(msa)} vs {row_of}"
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "Bio"))
        open(os.path.join(tmp, "Bio", "__init__.py"), "w").write("")
        open(os.path.join(tmp, "Bio", "SeqIO.py"), "w").write(SHIM)
        for mix, ref_key in (("TM", "Merlin"), ("TA", "AD169")):       # rules/genome_diff.smk:3-4: qry is always TB40E
            rm, rname, rlen = row_of[ref_key]
            qm, qname, qlen = row_of["TB40E"]
            rows = rows_from_msa(msa[rm], msa[qm], rname, qname, rlen, qlen)
            snps = os.path.join(tmp, mix + ".variants")
            open(snps, "w").write("\n".join(rows) + "\n")
            env = dict(os.environ, PYTHONPATH=tmp)
            vcf = subprocess.check_output([sys.executable, os.path.join(REF, "program", "mummer2vcf.py"), "-s", snps,
                                           "--output-header", "-n", "-g", os.path.join(REF, "ref", refs[ref_key])], env=env)
            lines = [ln for ln in vcf.decode().split("\n") if not ln.startswith("##fileDate")]   # keep the file reproducible
            out = os.path.join(OUT, f"{mix}.maskrepeat.variants.vcf.gz")
            with gzip.GzipFile(out, "wb", mtime=0) as fh:
                fh.write("\n".join(lines).encode())
            body = [ln for ln in lines if ln and not ln.startswith("#")]
            n_snp = sum(1 for ln in body if len(ln.split("\t")[3]) == 1 and len(ln.split("\t")[4]) == 1)
            print(mix, "rows", len(rows), "vcf records", len(body), "single-base SNP records", n_snp, "->", out)


if __name__ == "__main__":
    main()
