"""Loader for the packed genomes shipped with the package (quasimodo_b200/data/genomes/*.qmg, made by
tools/pack_refs.py from the reference's ref/*.fa and ref/Ecoli.NC_000913.fa.pac).  Host-side data
handling only."""
import os
import struct

import numpy as np

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "genomes")

# sample -> alignment reference, as rules/load_config.smk:20-23 of the reference
SAMPLE_REF = {"TM-0-1": "Merlin", "TM-1-1": "Merlin", "TM-1-10": "Merlin", "TM-1-50": "Merlin", "TM-1-0": "TB40E",
              "TA-1-0": "TB40E", "TA-1-1": "AD169", "TA-1-10": "AD169", "TA-1-50": "AD169", "TA-0-1": "AD169"}


class Genome:
    def __init__(self, names, lens, codes):
        self.names = names          # contig names
        self.lens = lens            # contig lengths
        self.codes = codes          # uint8 codes 0..3, contigs concatenated

    @property
    def total(self):
        return int(len(self.codes))

    def concat(self, other):
        return Genome(self.names + other.names, self.lens + other.lens, np.concatenate([self.codes, other.codes]))


def load(stem):
    path = os.path.join(DATA, stem + ".qmg")
    with open(path, "rb") as fh:
        if fh.read(4) != b"QMG1":
            raise ValueError(path + ": bad magic")
        (n,) = struct.unpack("<I", fh.read(4))
        names, lens = [], []
        for _ in range(n):
            (ln,) = struct.unpack("<I", fh.read(4))
            names.append(fh.read(ln).decode())
            lens.append(struct.unpack("<Q", fh.read(8))[0])
        (total,) = struct.unpack("<Q", fh.read(8))
        words = np.frombuffer(fh.read(), dtype="<u4")
    shifts = (2 * np.arange(16, dtype=np.uint32))
    codes = ((words[:, None] >> shifts[None, :]) & 3).astype(np.uint8).reshape(-1)[:total]
    return Genome(names, [int(x) for x in lens], codes)
