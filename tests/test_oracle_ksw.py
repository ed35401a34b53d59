"""Oracle self-checks (CPU): the C restatement of ksw_extend2 against the independent Python
restatement of SURVEY.md Appendix A.3, on random and adversarial tasks."""
import numpy as np
import pytest

from oracle import ksw_py, qmo_py


def mutate(rng, seq, sub=0.05, indel=0.02):
    out = []
    for b in seq:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(int(rng.integers(0, 4)))
        if rng.random() < sub:
            out.append(int(rng.integers(0, 4)))
        else:
            out.append(int(b))
    return np.array(out, dtype=np.uint8)


def random_task(rng):
    qlen = int(rng.integers(1, 130))
    q = rng.integers(0, 4, qlen).astype(np.uint8)
    kind = rng.integers(0, 4)
    if kind == 0:
        t = rng.integers(0, 4, int(rng.integers(0, 2 * qlen + 10))).astype(np.uint8)
    else:
        t = mutate(rng, q, sub=[0.0, 0.03, 0.12][kind - 1], indel=[0.0, 0.01, 0.05][kind - 1])
        t = np.concatenate([t, rng.integers(0, 4, int(rng.integers(0, qlen + 10))).astype(np.uint8)])
    if rng.random() < 0.1 and qlen > 2:
        q[rng.integers(0, qlen)] = 4
    if rng.random() < 0.05 and len(t) > 2:
        t[rng.integers(0, len(t))] = 4
    h0 = int(rng.integers(1, 160))
    w = int(rng.choice([1, 3, 5, 20, 100, 200]))
    eb = int(rng.choice([0, 5]))
    return q, t, h0, w, eb


def test_extend_c_vs_python_random():
    rng = np.random.default_rng(7)
    n_gs = 0
    for _ in range(1500):
        q, t, h0, w, eb = random_task(rng)
        c_res, c_cells = qmo_py.ksw_extend2(q, t, h0, w, eb)
        p_res, p_cells = ksw_py.ksw_extend2(list(q), list(t), h0, w, eb)
        assert c_res == p_res, (list(q), list(t), h0, w, eb)
        assert c_cells == p_cells
        n_gs += c_res[4] > 0
    assert n_gs > 100  # the to-end path is exercised


@pytest.mark.parametrize("q,t,h0,w,eb,expect", [
    # perfect 10-mer extension from h0=31: score 41 at (10,10); to-end score 41 at row 10
    ([0, 1, 2, 3, 0, 1, 2, 3, 0, 1], [0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 2], 31, 100, 5, (41, 10, 10, 10, 41, 0)),
    # empty target: nothing extends
    ([0, 1, 2], [], 31, 100, 5, (31, 0, 0, 0, -1, 0)),
    # all-N query: every cell scores -1, best stays h0
    ([4, 4, 4], [0, 1, 2, 3], 10, 100, 5, (10, 0, 0, 3, 7, 0)),
])
def test_extend_known_answers(q, t, h0, w, eb, expect):
    res, _ = qmo_py.ksw_extend2(q, t, h0, w, eb)
    py, _ = ksw_py.ksw_extend2(q, t, h0, w, eb)
    assert res == py
    assert res == expect


def test_global_simple():
    q = [0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3]
    s, cig = qmo_py.ksw_global2(q, q, 5)
    assert s == 12 and cig == [(0, 12)]
    t = q[:6] + [2, 2] + q[6:]           # 2-base deletion from the read's view
    s, cig = qmo_py.ksw_global2(q, t, 5)
    assert s == 12 - 8 and sum(l for op, l in cig if op in (0, 1)) == 12 and sum(l for op, l in cig if op in (0, 2)) == 14
    assert [op for op, _ in cig] == [0, 2, 0]


# ---- ksw_align2 (mate rescue's local alignment): C restatement vs the independent full-matrix Python one ----
def random_local_task(rng):
    qlen = int(rng.integers(1, 90))
    q = rng.integers(0, 4, qlen).astype(np.uint8)
    kind = rng.integers(0, 5)
    left = rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8)
    right = rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8)
    if kind == 0:
        t = rng.integers(0, 4, int(rng.integers(0, 150))).astype(np.uint8)
    elif kind == 4:      # two copies: the sub-optimal score is exercised
        t = np.concatenate([left, mutate(rng, q, 0.02, 0.0), right, mutate(rng, q, 0.06, 0.02), left])
    else:
        t = np.concatenate([left, mutate(rng, q, [0.0, 0.04, 0.12][kind - 1], [0.0, 0.02, 0.05][kind - 1]), right])
    if rng.random() < 0.15 and qlen > 2:
        q[rng.integers(0, qlen)] = 4
    if rng.random() < 0.1 and len(t) > 2:
        t[rng.integers(0, len(t))] = 4
    return q, t, int(rng.choice([0, 10, 19, 31]))


def test_align2_c_vs_python_random():
    rng = np.random.default_rng(11)
    n_sub = n_start = 0
    for _ in range(600):
        q, t, minsc = random_local_task(rng)
        c_res, c_cells = qmo_py.ksw_align2(q, t, minsc)
        p_res = ksw_py.ksw_align2(list(q), list(t), minsc)
        assert c_res == p_res, (list(q), list(t), minsc)
        n_sub += c_res[3] > 0
        n_start += c_res[6] >= 0
    assert n_sub > 30 and n_start > 200


def test_align2_known_answers():
    q = [0, 1, 2, 3, 0, 1, 2, 3, 2, 2, 1, 0]
    # exact copy at offset 3: score 12, ends inclusive, start recovered
    t = [3, 3, 3] + q + [0, 0]
    assert qmo_py.ksw_align2(q, t, 5)[0] == (12, 14, 11, -1, -1, 3, 0)
    # below minsc: no start
    assert qmo_py.ksw_align2(q, t, 13)[0] == (12, 14, 11, -1, -1, -1, -1)
    # empty target
    assert qmo_py.ksw_align2(q, [], 5)[0] == (0, -1, -1, -1, -1, -1, -1)
    # two exact copies: the first one wins, the second is the sub-optimal hit (more than `score` rows away)
    t2 = q + [3] * 20 + q
    r = qmo_py.ksw_align2(q, t2, 5)[0]
    assert r[0] == 12 and r[1] == 11 and r[3] == 12 and r[4] == 43 and (r[5], r[6]) == (0, 0)


# ---- ksw_global2 (CIGAR generation): the rolling-row C restatement vs the whole-matrix Python one ----
@pytest.mark.parametrize("scoring", [{}, dict(a=2, b=5, o_del=5, e_del=2, o_ins=7, e_ins=1), dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1),
                                     dict(a=1, b=9, o_del=0, e_del=1, o_ins=0, e_ins=3)])
def test_global_c_vs_python_random(scoring):
    """score and CIGAR, op for op: the tie rules of the traceback (diagonal over deletion over insertion; a gap continues only if
    extending beat opening strictly) decide where an indel inside a repeat lands, which is what downstream tools see"""
    rng = np.random.default_rng(42)
    opt = qmo_py.default_opt()
    for k, v in scoring.items():
        setattr(opt, k, v)
    kw = {k: getattr(opt, k) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins")}
    n_gapped = 0
    for it in range(400):
        qlen = int(rng.integers(1, 70))
        kind = it % 4
        if kind == 0:                   # low-complexity: many equally good placements of a gap
            unit = rng.integers(0, 4, int(rng.integers(1, 4))).astype(np.uint8)
            q = np.resize(unit, qlen)
        else:
            q = rng.integers(0, 4, qlen).astype(np.uint8)
        t = mutate(rng, q, [0.0, 0.03, 0.10, 0.05][kind], [0.08, 0.0, 0.04, 0.10][kind])
        if kind == 3 and rng.random() < 0.5:
            t = np.concatenate([t, rng.integers(0, 4, int(rng.integers(1, 6))).astype(np.uint8)])
        if rng.random() < 0.1:
            q = q.copy()
            q[rng.integers(0, qlen)] = 4
        if len(t) == 0:
            continue
        w = abs(len(t) - qlen) + int(rng.choice([3, 5, 12, 100]))
        got = qmo_py.ksw_global2(q, t, w, opt)
        want = ksw_py.ksw_global2(list(q), list(t), w, **kw)
        assert got == want, (it, list(q), list(t), w, got, want)
        assert sum(l for op, l in got[1] if op in (0, 1)) == qlen and sum(l for op, l in got[1] if op in (0, 2)) == len(t)
        n_gapped += any(op != 0 for op, _ in got[1])
    assert n_gapped > 100


def test_band_retry_of_the_extension_tasks_against_python():
    """the extension calls mem_chain2aln makes (the oracle's task log on reads with indels, band 2 so that retries happen): each task
    replayed through the independent ksw_py.ksw_extend2 with bwa's retry rule -- twice the band while the score still moves and the
    path came within a quarter of the band's edge, at most two tries -- gives the logged result, last band and cell count"""
    from quasimodo_b200 import workloads
    n = 150
    W = workloads.config5(n)
    codes, _, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, W.params.read_len, np.int32)
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    opt = qmo_py.default_opt()
    opt.w = 2
    lg = qmo_py.align_se(ref, codes, lens, opt=opt, want_log=True)["log"]
    seq = lg["seq"]
    n_retried = 0
    pick = np.random.default_rng(1).permutation(len(lg["tasks"]))[:160]
    for i in pick:
        t = lg["tasks"][i]
        q = [int(x) for x in seq[t["q_off"]:t["q_off"] + t["qlen"]]]
        tg = [int(x) for x in seq[t["t_off"]:t["t_off"] + t["tlen"]]]
        assert int(t["flags"]) & 1
        prev = int(t["h0"]) if int(t["flags"]) & 2 else -1
        cells = 0
        for attempt in range(2):
            w = int(t["w"]) << attempt
            res, c = ksw_py.ksw_extend2(q, tg, int(t["h0"]), w, int(t["end_bonus"]))
            cells += c
            if res[0] == prev or res[5] < (w >> 1) + (w >> 2):
                break
            prev = res[0]
        assert tuple(int(x) for x in lg["results"][i]) == res, (i, t)
        assert (int(lg["w_used"][i]), int(lg["cells"][i])) == (w, cells), (i, t)
        n_retried += w != int(t["w"])
    assert n_retried > 3, n_retried
