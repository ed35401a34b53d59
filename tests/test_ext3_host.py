"""CPU parity of the packed two-tasks-per-thread extension logic (quasimodo_b200/csrc/ext3_core.cuh): the statements the
CUDA kernel ext3_kernel runs, compiled for the host with the DPX instructions emulated (tests/ext3_host.cpp), against the
oracle's ksw_extend2 -- bit-exact, including the executed-cell count and the band retry.  The GPU tests then only have to
show that the kernel and this host build agree (tests/test_extend_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import qmo_py
from tests import extgen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIELDS = 8


@pytest.fixture(scope="module")
def host():
    bdir = os.path.join(HERE, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "libext3host.so")
    srcs = [os.path.join(HERE, "ext3_host.cpp"), os.path.join(ROOT, "quasimodo_b200", "csrc", "ext3_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-Wall", "-fPIC", "-shared", "-o", so, srcs[0]])
    return C.CDLL(so)


def oracle_results(pairs, h0s, ws, eb, flags, opt):
    out = []
    for (q, t), h0, w in zip(pairs, h0s, ws):
        cells, prev, res, wu = 0, (int(h0) if flags & 2 else -1), None, int(w)
        for a in range(2 if flags & 1 else 1):
            wu = int(w) << a
            res, c = qmo_py.ksw_extend2(q, t, int(h0), wu, eb, opt=opt)
            cells += c
            if res[0] == prev or res[5] < (wu >> 1) + (wu >> 2):
                break
            prev = res[0]
        out.append(tuple(res) + (wu, cells))
    return out


def run_host(host, pairs, h0s, ws, eb, flags=0, scoring=None, cap=256, order=None, narrow=None):
    opt = qmo_py.default_opt()
    for k, v in (scoring or {}).items():
        setattr(opt, k, v)
    n = len(pairs)
    order = np.arange(n) if order is None else np.asarray(order)
    chunks, q_off, t_off, off = [], [], [], 0
    for i in order:
        q, t = pairs[i]
        q_off.append(off); chunks.append(np.asarray(q, np.uint8)); off += len(q)
        t_off.append(off); chunks.append(np.asarray(t, np.uint8)); off += len(t)
    seq = np.concatenate(chunks + [np.zeros(1, np.uint8)])
    q_off, t_off = np.array(q_off, np.int64), np.array(t_off, np.int64)
    qlen = np.array([len(pairs[i][0]) for i in order], np.int32)
    tlen = np.array([len(pairs[i][1]) for i in order], np.int32)
    h0 = np.asarray(h0s, np.int32)[order].copy()
    w = np.asarray(ws, np.int32)[order].copy()
    fl = np.full(n, flags, np.uint32)
    sc = np.array([opt.a, opt.b, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.zdrop], np.int32)
    out = np.zeros((n, FIELDS), np.int32)
    bad = np.zeros(n, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if narrow is None:            # both plane layouts (the byte planes only hold schemes with a + b <= 16)
        n_run = run_host(host, pairs, h0s, ws, eb, flags, scoring, cap, order, narrow=0)
        if opt.a + opt.b <= 16:
            assert run_host(host, pairs, h0s, ws, eb, flags, scoring, cap, order, narrow=1) == n_run
        return n_run
    rc = host.ext3_host_run(p(sc), C.c_int(cap), C.c_int64(n), p(seq), p(q_off), p(t_off), p(qlen), p(tlen), p(h0), p(w), C.c_int(eb),
                            p(fl), p(out), p(bad), C.c_int(narrow))
    assert rc == 0
    want = oracle_results([pairs[i] for i in order], h0, w, eb, flags, opt)
    n_run, wrong = 0, []
    for k in range(n):
        if bad[k]:
            assert qlen[k] > cap or qlen[k] > 256 or h0[k] + qlen[k] * opt.a > 255
            continue
        n_run += 1
        if tuple(int(x) for x in out[k]) != want[k]:
            wrong.append((int(order[k]), tuple(int(x) for x in out[k]), want[k], int(qlen[k]), int(tlen[k]), int(h0[k]), int(w[k])))
    assert not wrong, f"{len(wrong)} of {n_run} tasks differ; first: {wrong[:3]}"
    return n_run


def by_qlen(pairs):
    return np.argsort([len(q) for q, _ in pairs], kind="stable")


@pytest.mark.parametrize("eb", [0, 5])
def test_random_tasks_in_random_pairs(host, eb):
    """partners of very different shape: the half-word blends before and after the common columns carry most cells"""
    rng = np.random.default_rng(1234 + eb)
    pairs, h0s, ws = extgen.random_tasks(rng, 3000)
    h0s = np.minimum(h0s, 255 - np.array([len(q) for q, _ in pairs]))
    h0s = np.maximum(h0s, 1)
    assert run_host(host, pairs, h0s, ws, eb) > 2500


def test_random_tasks_sorted_by_query_length(host):
    """what the pipeline feeds the kernel: partners of equal query length"""
    rng = np.random.default_rng(7)
    pairs, h0s, ws = extgen.random_tasks(rng, 3000)
    h0s = np.maximum(np.minimum(h0s, 255 - np.array([len(q) for q, _ in pairs])), 1)
    assert run_host(host, pairs, h0s, ws, 5, order=by_qlen(pairs)) > 2500


@pytest.mark.parametrize("flags", [1, 3])
def test_band_retry(host, flags):
    rng = np.random.default_rng(4321)
    pairs, h0s, ws = extgen.random_tasks(rng, 2000)
    ws = np.where(ws > 50, 10, ws)        # small bands so that the 2w retry actually triggers
    h0s = np.maximum(np.minimum(h0s, 255 - np.array([len(q) for q, _ in pairs])), 1)
    assert run_host(host, pairs, h0s, ws, 5, flags=flags) > 1500
    assert run_host(host, pairs, h0s, ws, 5, flags=flags, order=by_qlen(pairs)) > 1500


def test_adversarial(host):
    pairs, h0s, ws = extgen.adversarial_tasks()
    for eb in (0, 5):
        assert run_host(host, pairs, h0s, ws, eb) > 40
        assert run_host(host, pairs, h0s, ws, eb, order=np.arange(len(pairs))[::-1]) > 40
        assert run_host(host, pairs, h0s, ws, eb, cap=64) > 10      # a class too small for most of them: refusals, no overrun


@pytest.mark.parametrize("scoring", [dict(a=2, b=5), dict(o_del=5, e_del=2, o_ins=7, e_ins=1), dict(a=3, b=4, o_del=4, e_del=3, o_ins=9, e_ins=2),
                                     dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1), dict(a=1, b=9, o_del=0, e_del=1, o_ins=0, e_ins=3)])
def test_other_scoring_schemes(host, scoring):
    rng = np.random.default_rng(5)
    pairs, h0s, ws = extgen.random_tasks(rng, 2500, max_qlen=70)
    a = scoring.get("a", 1)
    h0s = np.maximum(np.minimum(rng.integers(1, 250, len(h0s)), 255 - a * np.array([len(q) for q, _ in pairs])), 1)
    assert run_host(host, pairs, h0s, ws, 5, scoring=scoring) > 2000
    assert run_host(host, pairs, h0s, ws, 5, scoring=scoring, order=by_qlen(pairs)) > 2000


def test_read_like_tasks_with_n(host):
    """pipeline-shaped tasks (query 20..119, target = query + max gap, one mismatch up front) with N in query and target"""
    rng = np.random.default_rng(99)
    pairs, h0s, ws = [], [], []
    for _ in range(1500):
        ql = int(rng.integers(20, 120))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = np.concatenate([q, rng.integers(0, 4, ql - 5).astype(np.uint8)])
        t[0] = (t[0] + 1) & 3
        if rng.random() < 0.5:
            q[rng.integers(0, ql, 2)] = 4
        if rng.random() < 0.3:
            t[rng.integers(0, len(t), 3)] = 4
        if rng.random() < 0.2:
            t = np.delete(t, rng.integers(5, ql - 5))
        pairs.append((q, t)); h0s.append(int(rng.integers(31, 130))); ws.append(100)
    assert run_host(host, pairs, np.array(h0s), np.array(ws), 5, flags=3, order=by_qlen(pairs)) == 1500
