"""Multi-GPU decomposition of one sample: contiguous pair-index ranges per rank, the sample's
insert-size prefix replicated on every rank, one integer all-reduce of the count tensor.  Host-side logic only;
works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""


def shard_range(n_pairs, rank, world):
    """pairs [lo, hi) of rank `rank`: contiguous, sizes differ by at most one, union = [0, n_pairs)"""
    base, extra = divmod(int(n_pairs), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def prefix_range(n_pairs, prefix_pairs):
    """the designated insert-size prefix of the sample: pairs [0, min(n, prefix_pairs))"""
    return 0, min(int(n_pairs), int(prefix_pairs))


def needs_prefix(rank_lo, rank_hi, n_pairs, prefix_pairs):
    """True when the shard does not itself start with the whole prefix (then qm_sample_estimate_pestat must be
    called with the sample's first pairs before the shard's own pairs are added)"""
    return not (rank_lo == 0 and rank_hi - rank_lo >= min(n_pairs, prefix_pairs))


def allreduce_counts(counts, group=None):
    """sum the per-rank int32 count tensors in place (integer sum: order independent, bit-exact for any world size)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts
