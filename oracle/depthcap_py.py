"""ORACLE (test infrastructure) -- a second restatement of htslib's pileup depth cap (`bcftools mpileup -d N`, SURVEY.md A.8 /
8f-2), written as a literal replay of the iterator's linked list (htslib 1.9 sam.c bam_plp_push / bam_plp_next / bam_plp_auto as
published; htslib is not in the image), to be diffed against the heap-based C restatement in oracle/qmo_pileup.c.

State of the iterator: a list of the reads pushed and not yet freed (plus ONE spare node at the tail that the next read is
copied into: it counts), the position (tid, pos) being assembled, and the start of the last read pushed.  bam_plp_auto hands the
iterator one read whenever it has caught up with the last read's start; bam_plp_next frees the reads that ended at or before the
position it assembles and then moves on -- to the list's first read if that starts further right, else one base."""
import numpy as np

from oracle import sort_py


def _admitted(a, min_mapq, count_orphans):
    f = int(a["flag"])
    return not (f & (0x4 | 0x100 | 0x200 | 0x400)) and int(a["n_cigar"]) not in (0, 255) and int(a["mapq"]) >= min_mapq \
        and not ((f & 0x1) and not (f & 0x2) and not count_orphans)


def depth_cap(alns, max_depth, min_mapq=0, count_orphans=False):
    """-> bool mask over alns: reads the iterator keeps (False: dropped by the cap, or never admitted)"""
    keep = np.zeros(len(alns), dtype=bool)
    nodes = []                                   # [tid, beg, end] of linked reads, in arrival order
    it_tid, it_pos = 0, 0
    last_tid, last_beg = -1, -1                  # iter->max_tid / max_pos
    for gi in sort_py.sort_perm(alns):
        a = alns[int(gi)]
        if not _admitted(a, min_mapq, count_orphans):
            continue                             # the read function skips it: the iterator never sees it
        tid, beg = int(a["rid"]), int(a["pos"])
        rlen = sum(int(c) >> 4 for c in a["cigar"][:int(a["n_cigar"])] if (int(c) & 15) in (0, 2, 3, 7, 8))
        end = beg + max(rlen, 1)
        # --- bam_plp_push ---
        if not (it_tid == tid and it_pos == beg and len(nodes) + 1 > max_depth):
            last_tid, last_beg = tid, beg
            if end > it_pos or tid > it_tid:
                nodes.append((tid, beg, end))
                keep[int(gi)] = True
        # --- bam_plp_next until it has nothing beyond its position ---
        while last_tid > it_tid or (last_tid == it_tid and last_beg > it_pos):
            nodes = [p for p in nodes if not (p[0] < it_tid or (p[0] == it_tid and p[2] <= it_pos))]
            if nodes and it_tid < nodes[0][0]:
                it_tid, it_pos = nodes[0][0], nodes[0][1]
            elif nodes and it_pos < nodes[0][1]:
                it_pos = nodes[0][1]
            else:
                it_pos += 1
    return keep
