#!/usr/bin/env python3
"""Pack the reference genomes bundled with QuasiModo into this repo's .qmg format.

Run ONCE in the build container (where /root/reference is mounted); the GPU box
has no /root/reference, so the packed genomes travel with the repo as data.

Sources (reference repo, read-only):
  ref/Merlin.BAC.fa, ref/TB40E.GFP.fa, ref/AD169.BAC.fa, ref/Phix.fa   (FASTA)
  ref/Ecoli.NC_000913.fa.pac + .ann   (the FASTA itself is a missing blob; the
      genome is recovered from bwa's 2-bit .pac: base i sits in byte i>>2 at
      shift ((~i)&3)<<1 -- SURVEY.md section 0.6)

.qmg layout (little-endian), deliberately NOT the bwa .pac layout:
  char[4]  "QMG1"
  u32      n_contigs
  per contig: u32 name_len, name bytes (no NUL), u64 length
  u64      total_len
  u32[ceil(total_len/16)]  bases, 2 bit each, base i in word i>>4 at bit 2*(i&15)
  codes: A=0 C=1 G=2 T=3.  All bundled genomes are pure ACGT (verified here).
"""
import os
import struct
import sys

import numpy as np

REF = "/root/reference/ref"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "quasimodo_b200", "data", "genomes")


def read_fasta(path):
    contigs = []
    name, chunks = None, []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                if name is not None:
                    contigs.append((name, "".join(chunks)))
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.upper())
    contigs.append((name, "".join(chunks)))
    return contigs


def codes_from_seq(seq):
    a = np.frombuffer(seq.encode(), dtype=np.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    for ch, c in zip(b"ACGT", range(4)):
        lut[ch] = c
    codes = lut[a]
    if (codes == 255).any():
        raise SystemExit("ambiguous base found; .qmg is 2-bit only")
    return codes


def codes_from_pac(pac_path, l_pac):
    raw = np.fromfile(pac_path, dtype=np.uint8)
    i = np.arange(l_pac, dtype=np.int64)
    return ((raw[i >> 2] >> (((~i) & 3) << 1)) & 3).astype(np.uint8)


def write_qmg(path, contigs):
    """contigs: list of (name, uint8 code array)"""
    allc = np.concatenate([c for _, c in contigs])
    n = len(allc)
    pad = (-n) % 16
    padded = np.concatenate([allc, np.zeros(pad, dtype=np.uint8)]).astype(np.uint32).reshape(-1, 16)
    shifts = (2 * np.arange(16, dtype=np.uint32))
    words = (padded << shifts).sum(axis=1).astype("<u4")
    with open(path, "wb") as fh:
        fh.write(b"QMG1")
        fh.write(struct.pack("<I", len(contigs)))
        for name, c in contigs:
            nb = name.encode()
            fh.write(struct.pack("<I", len(nb)))
            fh.write(nb)
            fh.write(struct.pack("<Q", len(c)))
        fh.write(struct.pack("<Q", n))
        fh.write(words.tobytes())
    print(f"{path}: {len(contigs)} contig(s), {n} bp, {os.path.getsize(path)} B")


def main():
    os.makedirs(OUT, exist_ok=True)
    for stem, fa in [("Merlin", "Merlin.BAC.fa"), ("TB40E", "TB40E.GFP.fa"),
                     ("AD169", "AD169.BAC.fa"), ("Phix", "Phix.fa")]:
        contigs = [(n, codes_from_seq(s)) for n, s in read_fasta(os.path.join(REF, fa))]
        # cross-check against bwa's own .pac (pins the 2-bit code assignment)
        l_pac = sum(len(c) for _, c in contigs)
        pac = codes_from_pac(os.path.join(REF, fa + ".pac"), l_pac)
        assert (pac == np.concatenate([c for _, c in contigs])).all(), fa
        write_qmg(os.path.join(OUT, stem + ".qmg"), contigs)
    # E. coli: recover from .pac
    with open(os.path.join(REF, "Ecoli.NC_000913.fa.ann")) as fh:
        l_pac = int(fh.readline().split()[0])
        name = fh.readline().split()[1]
    codes = codes_from_pac(os.path.join(REF, "Ecoli.NC_000913.fa.pac"), l_pac)
    write_qmg(os.path.join(OUT, "Ecoli.qmg"), [(name, codes)])


if __name__ == "__main__":
    sys.exit(main())
