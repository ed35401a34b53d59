"""CPU restatement of the reference's record order and BAM index arithmetic.  TEST INFRASTRUCTURE ONLY (imported by
tests/ alone; never by the product package).

`samtools sort` (rules/bwa.smk:17 of the reference; samtools 1.9 bam_sort.c, bam1_lt) orders records by
    (uint64)tid << 32 | (pos + 1) << 1 | is_reverse
with ties kept in input order; tid = -1 wraps to the top, so unplaced records come last.
samtools is not in the image and not vendored (conda pin config/conda_env.yaml:11): PARITY UNPINNED -- the formula
above is the published comparator, restated; the reference holds no golden BAM.

reg2bin / the 16 kb linear index follow the SAM specification section 5."""
import numpy as np


def samtools_keys(alns):
    tid = alns["rid"].astype(np.int64).astype(np.uint64) & np.uint64(0xffffffff)      # (uint64)(uint32)tid
    pos1 = (alns["pos"].astype(np.int64) + 1).astype(np.uint64)
    rev = ((alns["flag"] & 0x10) != 0).astype(np.uint64)
    return (tid << np.uint64(32)) | (pos1 << np.uint64(1)) | rev


def sort_perm(alns):
    """perm[i] = input index of the record at sorted position i"""
    return np.argsort(samtools_keys(alns), kind="stable").astype(np.uint32)


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def reg2bins(beg, end):
    """every bin that may hold records overlapping [beg, end)"""
    end -= 1
    bins = [0]
    for shift, off in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(off + (beg >> shift), off + (end >> shift) + 1))
    return bins
