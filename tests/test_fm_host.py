"""CPU parity of the device FM-index seeding (quasimodo_b200/csrc/fm_core.cuh compiled for the host, tests/fm_host.cpp) against
the oracle's restatement of bwa's seeding (oracle/qmo_fm.c): the same seeds in the same order, read by read -- on simulated
reads of the BASELINE mixtures (both strands, substitutions, indels, N), on a genome with diverged repeats, with and without
the third seeding round."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import qmo_py
from quasimodo_b200 import genomes, workloads

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MAX_SEEDS = 64


@pytest.fixture(scope="module")
def host():
    bdir = os.path.join(HERE, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libfmhost.so")
    srcs = [os.path.join(HERE, "fm_host.cpp"), os.path.join(ROOT, "quasimodo_b200", "csrc", "fm_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", so, srcs[0]])
    return C.CDLL(so)


def host_seeds(host, fm, lens_c, reads, lens, k, max_occ, max_mem_intv):
    b, s = np.frombuffer(fm.bwt_bytes(), np.uint8), np.frombuffer(fm.sa_bytes(), np.uint8)
    lens_c = np.ascontiguousarray(lens_c, np.int64)
    off = np.concatenate([[0], np.cumsum(lens_c)[:-1]]).astype(np.int64)
    reads = np.ascontiguousarray(reads, np.uint8)
    n, stride = reads.shape
    lens = np.ascontiguousarray(lens, np.int32)
    out = np.zeros((n, MAX_SEEDS, 3), np.int64)
    n_out = np.zeros(n, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = host.fm_host_seeds(p(b), C.c_int64(len(b)), p(s), C.c_int64(len(s)), C.c_int(len(lens_c)), p(off), p(lens_c), C.c_int64(int(lens_c.sum())),
                            C.c_int(k), C.c_int(max_occ), C.c_int(max_mem_intv), C.c_int64(n), p(reads), C.c_int(stride), p(lens),
                            C.c_int(MAX_SEEDS), p(out), p(n_out))
    assert rc == 0
    return out, n_out


def compare(host, G_codes, G_lens, reads, lens, k=31, max_mem_intv=20):
    fm = qmo_py.FmIndex(codes=G_codes)
    ref = qmo_py.Ref(G_codes, G_lens, k=k)
    opt = qmo_py.default_opt()
    opt.min_seed_len = k
    got, n_got = host_seeds(host, fm, G_lens, reads, lens, k, opt.max_occ, max_mem_intv)
    total = 0
    for r in range(len(reads)):
        want = fm.seeds(ref, reads[r, :lens[r]], opt, max_mem_intv=max_mem_intv, max_seeds=MAX_SEEDS)
        assert n_got[r] == len(want) and np.array_equal(got[r, :n_got[r]], want), (r, got[r, :n_got[r]].tolist(), want.tolist())
        total += len(want)
    return total


@pytest.mark.parametrize("cfg,n", [("cfg2", 1500), ("cfg5", 800), ("cfg3", 1000)])
def test_seeds_equal_the_oracles(host, cfg, n):
    W = {"cfg2": lambda: workloads.config2(4, n), "cfg5": lambda: workloads.config5(n), "cfg3": lambda: workloads.config3(n)}[cfg]()
    codes, _ = qmo_py.simulate_pairs(W, 0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    if cfg == "cfg3":                      # E. coli is 4.6 Mb: its suffix array costs 15 s; Merlin | PhiX carry the multi-contig logic
        G = genomes.load("Merlin").concat(genomes.load("Phix"))
    else:
        G = W.ref
    assert compare(host, G.codes, G.lens, codes, lens) > n
    assert compare(host, G.codes, G.lens, codes[:300], lens[:300], max_mem_intv=0) > 100


def test_seeds_on_a_repetitive_genome(host):
    """tandem and dispersed copies with a few substitutions: many occurrences per interval, re-seeding, the occurrence cap"""
    rng = np.random.default_rng(21)
    unit = rng.integers(0, 4, 300).astype(np.uint8)
    parts = [rng.integers(0, 4, 1000).astype(np.uint8)]
    for c in range(40):
        u = unit.copy()
        for _ in range(int(rng.integers(0, 4))):
            u[rng.integers(0, 300)] = rng.integers(0, 4)
        parts += [u, rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8)]
    g = np.concatenate(parts)
    reads = np.full((400, 150), 4, np.uint8)
    lens = np.full(400, 150, np.int32)
    for r in range(400):
        p = int(rng.integers(0, len(g) - 150))
        q = g[p:p + 150].copy()
        if r % 2:
            q = (3 - q[::-1]).astype(np.uint8)
        for _ in range(int(rng.integers(0, 3))):
            q[rng.integers(0, 150)] = rng.integers(0, 5)
        reads[r] = q
    assert compare(host, g, [len(g)], reads, lens, k=19) > 400 * 10
    assert compare(host, g, [len(g)], reads[:100], lens[:100], k=19, max_mem_intv=0) > 100
