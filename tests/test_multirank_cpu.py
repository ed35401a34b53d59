"""world_size-2 gloo test (CPU) of the N>1 host logic: shard ranges, replicated insert-size prefix, integer
all-reduce of the count tensors.  The per-shard compute is done by the oracle here (there is no GPU in the build
box); on the GPU box the same decomposition is exercised by tests/test_sample_gpu.py::test_sample_sharded_equals_whole
and by `bench.py --gpus N`."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import qmo_py
    from quasimodo_b200 import sharding, workloads
    W = workloads.config2(4, n)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    lo, hi = sharding.shard_range(n, rank, world)
    codes, quals, _, _ = W.simulate_host(lo, hi - lo)              # index-addressable: each rank simulates its own shard
    lens = np.full(2 * (hi - lo), 150, np.int32)
    prefix = None
    if sharding.needs_prefix(lo, hi, n, qmo_py.PESTAT_PAIRS):
        plo, phi = sharding.prefix_range(n, qmo_py.PESTAT_PAIRS)
        pc, _, _, _ = W.simulate_host(plo, phi - plo)
        prefix = (pc, np.full(2 * (phi - plo), 150, np.int32))
    _, counts, _, pes = qmo_py.run_sample(ref, codes, quals, lens, pair_id0=lo, prefix=prefix)
    t = torch.from_numpy(counts)
    sharding.allreduce_counts(t)
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), t.numpy())
        np.save(os.path.join(out_dir, "pes.npy"), pes)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import qmo_py
    from quasimodo_b200 import sharding, workloads
    n = 3001                                                       # odd: ragged shards
    assert sharding.shard_range(n, 0, 2) == (0, 1501) and sharding.shard_range(n, 1, 2) == (1501, 3001)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    W = workloads.config2(4, n)
    codes, quals, _, _ = W.simulate_host(0, n)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    _, want, _, pes = qmo_py.run_sample(ref, codes, quals, np.full(2 * n, 150, np.int32))
    assert np.array_equal(np.load(tmp_path / "sum.npy"), want)
    assert np.load(tmp_path / "pes.npy").tobytes() == pes.tobytes()


def test_shard_ranges_cover():
    sys.path.insert(0, ROOT)
    from quasimodo_b200 import sharding
    for n in (0, 1, 7, 8, 50_000_000):
        for w in (1, 2, 4, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
