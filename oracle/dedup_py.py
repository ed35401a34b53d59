"""CPU restatement of `picard MarkDuplicates` as the reference runs it (rules/rmdup.smk:13-16: REMOVE_DUPLICATES=true,
everything else default).  TEST INFRASTRUCTURE ONLY (imported by tests/ alone; never by the product package).

picard (conda pin config/conda_env.yaml) is not in the image and not vendored: PARITY UNPINNED -- this follows the
published rule: fully placed pairs are compared by {contig, unclipped 5' coordinate, strand} of both
ends; the pair with the largest sum of base qualities >= 15 stays, ties go to the first in the file; a read whose mate
is unplaced is a duplicate whenever any end of a fully placed pair has its {contig, unclipped 5' coordinate, strand},
otherwise the best such read stays; pairs without a placed end are never duplicates."""
import numpy as np


def _end(a):
    """(contig, unclipped 5' coordinate, strand) of a placed record"""
    nc = int(a["n_cigar"])
    cig = [(int(c) & 0xf, int(c) >> 4) for c in a["cigar"][:nc]]
    rlen = sum(l for op, l in cig if op in (0, 2))
    lead = cig[0][1] if cig and cig[0][0] == 4 else 0
    trail = cig[-1][1] if len(cig) > 1 and cig[-1][0] == 4 else 0
    rev = bool(a["flag"] & 0x10)
    coord = int(a["pos"]) + rlen - 1 + trail if rev else int(a["pos"]) - lead
    return (int(a["rid"]), coord, int(rev))


def _placed(a):
    return not (a["flag"] & 0x4) and a["rid"] >= 0 and a["n_cigar"] not in (0, 255)


def mark_duplicates(alns, quals, lens):
    """alns: records 2 per pair in input order; quals [2n, stride]; -> bool array per pair (True = duplicate)"""
    n = len(alns) // 2
    score1 = np.array([int(q[:l][q[:l] >= 15].sum()) for q, l in zip(quals, lens)], dtype=np.int64)
    pairs, frags, pair_ends = {}, {}, set()
    for i in range(n):
        a, b = alns[2 * i], alns[2 * i + 1]
        pa, pb = _placed(a), _placed(b)
        if pa and pb:
            e = sorted([_end(a), _end(b)])
            pairs.setdefault((e[0], e[1]), []).append(i)
            pair_ends.update(e)
        elif pa or pb:
            frags.setdefault(_end(a if pa else b), []).append(i)
    dup = np.zeros(n, dtype=bool)
    for members in pairs.values():
        sc = [score1[2 * i] + score1[2 * i + 1] for i in members]
        best = members[int(np.argmax(sc))]              # argmax returns the first maximum: earliest in the file
        for i in members:
            dup[i] = i != best
    for key, members in frags.items():
        if key in pair_ends:
            dup[members] = True
            continue
        sc = [score1[2 * i] if _placed(alns[2 * i]) else score1[2 * i + 1] for i in members]
        best = members[int(np.argmax(sc))]
        for i in members:
            dup[i] = i != best
    return dup
