"""ORACLE (test infrastructure) -- the duplicate rule of `picard MarkDuplicates` (rules/rmdup.smk:13-16) a second
time, as picard itself works: one sorted table of read ends, runs of equal keys, the best of a run survives -- where
oracle/dedup_py.py groups with dictionaries.  Vectorised with numpy; to be diffed against dedup_py.mark_duplicates."""
import numpy as np


def _ends(alns):
    """per record: placed?, contig, unclipped 5' coordinate, strand"""
    n = len(alns)
    ops, lns = alns["cigar"] & 0xf, (alns["cigar"] >> 4).astype(np.int64)
    nc = alns["n_cigar"].astype(np.int64)
    valid = np.arange(alns["cigar"].shape[1])[None, :] < np.where(nc == 255, 0, nc)[:, None]
    span = np.where(valid & ((ops == 0) | (ops == 2)), lns, 0).sum(1)
    lead = np.where(valid[:, 0] & (ops[:, 0] == 4), lns[:, 0], 0)
    last = np.clip(nc - 1, 0, alns["cigar"].shape[1] - 1)
    rows = np.arange(n)
    trail = np.where((nc > 1) & (nc != 255) & (ops[rows, last] == 4), lns[rows, last], 0)
    rev = (alns["flag"] & 0x10) != 0
    coord = np.where(rev, alns["pos"].astype(np.int64) + span - 1 + trail, alns["pos"].astype(np.int64) - lead)
    placed = ((alns["flag"] & 0x4) == 0) & (alns["rid"] >= 0) & (nc != 0) & (nc != 255)
    return placed, alns["rid"].astype(np.int64), coord, rev.astype(np.int64)


def mark_duplicates(alns, quals, lens):
    n = len(alns) // 2
    placed, rid, coord, rev = _ends(alns)
    cols = np.arange(quals.shape[1])[None, :] < np.asarray(lens)[:, None]
    q = np.where(cols & (quals >= 15), quals, 0).sum(1).astype(np.int64)
    end = (rid << 40) | ((coord + (1 << 20)) << 1) | rev             # one integer per read end
    e0, e1, p0, p1 = end[0::2], end[1::2], placed[0::2], placed[1::2]
    dup = np.zeros(n, dtype=bool)
    # fully placed pairs: key = the two ends in ascending order; within a run of equal keys the best score stays, earliest on ties
    full = np.flatnonzero(p0 & p1)
    lo, hi = np.minimum(e0[full], e1[full]), np.maximum(e0[full], e1[full])
    score = q[0::2][full] + q[1::2][full]
    order = np.lexsort((full, -score, hi, lo))
    newrun = np.ones(len(order), dtype=bool)
    newrun[1:] = (lo[order][1:] != lo[order][:-1]) | (hi[order][1:] != hi[order][:-1])
    dup[full[order[~newrun]]] = True
    # pairs with one placed end: duplicates wherever a fully placed pair has an end, otherwise the same run rule among themselves
    half = np.flatnonzero(p0 ^ p1)
    key = np.where(p0[half], e0[half], e1[half])
    score = np.where(p0[half], q[0::2][half], q[1::2][half])
    taken = np.isin(key, np.concatenate([e0[full], e1[full]]))
    dup[half[taken]] = True
    rest, key, score = half[~taken], key[~taken], score[~taken]
    order = np.lexsort((rest, -score, key))
    newrun = np.ones(len(order), dtype=bool)
    newrun[1:] = key[order][1:] != key[order][:-1]
    dup[rest[order[~newrun]]] = True
    return dup
