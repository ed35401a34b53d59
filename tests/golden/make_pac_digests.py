#!/usr/bin/env python3
"""Golden digests of the reference's OWN bwa index files (ref/*.pac, *.ann, *.fai under /root/reference), generated in the
build container where the reference is mounted; committed as tests/golden/pac_digests.json.  They pin the one part of the
bwa side of the path the reference ships artefacts for ("only ref/*.{pac,bwt,sa,ann,amb} pin index
construction"): the packed genomes this repo aligns against must be base-for-base what bwa's index holds.

.pac layout (bwa bntseq.c): base i in byte i>>2 at shift ((~i)&3)<<1; the last byte holds l_pac % 4 (after an extra zero
byte when l_pac % 4 == 0); .ann line 1 = l_pac n_seqs seed, then per contig "gi name comment" / "offset len n_ambs"."""
import hashlib
import json
import os

import numpy as np

REF = "/root/reference/ref"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pac_digests.json")
STEMS = {"Merlin": "Merlin.BAC.fa", "TB40E": "TB40E.GFP.fa", "AD169": "AD169.BAC.fa", "Phix": "Phix.fa", "Ecoli": "Ecoli.NC_000913.fa"}


def main():
    out = {}
    for stem, base in STEMS.items():
        ann = open(os.path.join(REF, base + ".ann")).read().split("\n")
        l_pac, n_seqs = int(ann[0].split()[0]), int(ann[0].split()[1])
        names, lens = [], []
        for c in range(n_seqs):
            names.append(ann[1 + 2 * c].split()[1])
            lens.append(int(ann[2 + 2 * c].split()[1]))
        pac = np.fromfile(os.path.join(REF, base + ".pac"), dtype=np.uint8)
        i = np.arange(l_pac, dtype=np.int64)
        codes = ((pac[i >> 2] >> (((~i) & 3) << 1)) & 3).astype(np.uint8)
        out[stem] = {"file": base + ".pac", "l_pac": l_pac, "names": names, "lens": lens, "sha256_codes": hashlib.sha256(codes.tobytes()).hexdigest()}
    json.dump(out, open(OUT, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
