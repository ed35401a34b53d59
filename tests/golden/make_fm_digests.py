"""Digests of bwa's own FM-index files shipped in the reference (ref/*.bwt, ref/*.sa): the pin of the index construction
restated in oracle/qmo_fm.c.  Run in the build container (reads /root/reference/ref); writes tests/golden/fm_digests.json."""
import hashlib
import json
import os

REF = "/root/reference/ref"
FILES = {"Merlin": "Merlin.BAC.fa", "TB40E": "TB40E.GFP.fa", "AD169": "AD169.BAC.fa", "Phix": "Phix.fa", "Ecoli": "Ecoli.NC_000913.fa"}

out = {}
for stem, f in FILES.items():
    e = {}
    for ext in ("bwt", "sa"):
        p = os.path.join(REF, f + "." + ext)
        if os.path.exists(p):
            b = open(p, "rb").read()
            e[ext] = {"bytes": len(b), "sha256": hashlib.sha256(b).hexdigest()}
    out[stem] = e
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fm_digests.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
