"""ORACLE (test infrastructure) -- a second restatement of bwa-mem's insert-size model (bwamem_pair.c mem_pestat),
written from the published description with numpy, to be diffed against the C restatement in oracle/qmo_mem.c (bwa is not in the
image: neither can be pinned, two independent ones can at least be held against each other).

Input: the per-read region lists of the single-end stage (best region first), the length of the packed reference.
Output: for the four orientations (FF, FR, RF, RR) low / high / failed / avg / std."""
import math

import numpy as np


def _second_best(regs, n, mask_level, floor):
    """score of the first lower region that overlaps the best one by at least mask_level of the shorter of the two (bwa's
    cal_sub), else `floor` (= min_seed_len * a)"""
    b = regs[0]
    for j in range(1, n):
        r = regs[j]
        lo, hi = max(int(r["qb"]), int(b["qb"])), min(int(r["qe"]), int(b["qe"]))
        if hi > lo:
            shorter = min(int(r["qe"]) - int(r["qb"]), int(b["qe"]) - int(b["qb"]))
            if hi - lo >= shorter * mask_level:
                return int(r["score"])
    return floor


def orientation_and_distance(l_pac, b1, b2):
    """bwa's mem_infer_dir: mate 2's start mirrored onto mate 1's strand; 0 FF, 1 FR, 2 RF, 3 RR"""
    r1, r2 = b1 >= l_pac, b2 >= l_pac
    p2 = b2 if r1 == r2 else 2 * l_pac - 1 - b2
    return (0 if r1 == r2 else 1) ^ (0 if p2 > b1 else 3), abs(p2 - b1)


def pestat(l_pac, regs, n_regs, a=1, min_seed_len=31, mask_level=0.5, max_ins=10000):
    sizes = [[], [], [], []]
    for p in range(len(n_regs) // 2):
        n0, n1 = int(n_regs[2 * p]), int(n_regs[2 * p + 1])
        if n0 == 0 or n1 == 0:
            continue
        r0, r1 = regs[2 * p], regs[2 * p + 1]
        if _second_best(r0, n0, mask_level, min_seed_len * a) > 0.8 * int(r0[0]["score"]):
            continue
        if _second_best(r1, n1, mask_level, min_seed_len * a) > 0.8 * int(r1[0]["score"]):
            continue
        if int(r0[0]["rid"]) != int(r1[0]["rid"]):
            continue
        d, dist = orientation_and_distance(l_pac, int(r0[0]["rb"]), int(r1[0]["rb"]))
        if 0 < dist <= max_ins:
            sizes[d].append(dist)
    out = []
    for d in range(4):
        q = np.sort(np.array(sizes[d], dtype=np.int64))
        if len(q) < 10:
            out.append(dict(low=0, high=0, failed=1, avg=0.0, std=0.0))
            continue
        p25, p50, p75 = (int(q[int(f * len(q) + .499)]) for f in (.25, .50, .75))
        lo = max(1, int(p25 - 2.0 * (p75 - p25) + .499))
        hi = int(p75 + 2.0 * (p75 - p25) + .499)
        inside = [int(v) for v in q if lo <= v <= hi]
        avg = 0.0
        for v in inside:                                   # summed in sorted order, as the reference loop does
            avg += v
        avg /= len(inside)
        var = 0.0
        for v in inside:
            var += (v - avg) * (v - avg)
        std = math.sqrt(var / len(inside))
        lo = int(p25 - 3.0 * (p75 - p25) + .499)
        hi = int(p75 + 3.0 * (p75 - p25) + .499)
        if lo > avg - 4.0 * std:
            lo = int(avg - 4.0 * std + .499)
        if hi < avg + 4.0 * std:
            hi = int(avg + 4.0 * std + .499)
        out.append(dict(low=max(1, lo), high=hi, failed=0, avg=avg, std=std))
    most = max(len(s) for s in sizes)
    for d in range(4):
        if not out[d]["failed"] and len(sizes[d]) < most * 0.05:
            out[d]["failed"] = 1
    return out, [len(s) for s in sizes]
