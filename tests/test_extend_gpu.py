"""GPU parity (bit-exact): the CUDA batched extension kernel, called through the C-ABI, against the CPU
oracle's ksw_extend2 on the same tasks.  Integer work => every field must be identical."""
import numpy as np
import pytest

from tests import extgen

pytestmark = pytest.mark.gpu

FIELDS = ("score", "qle", "tle", "gtle", "gscore", "max_off")


@pytest.fixture(scope="module")
def ctx():
    from quasimodo_b200 import Context
    c = Context(0)
    yield c
    c.close()


def oracle_results(pairs, h0s, ws, eb, retry=False, prev_h0=False, opt=None):
    from oracle import qmo_py
    out = []
    for (q, t), h0, w in zip(pairs, h0s, ws):
        cells, prev, res, wu = 0, (int(h0) if prev_h0 else -1), None, int(w)
        for a in range(2 if retry else 1):
            wu = int(w) << a
            res, c = qmo_py.ksw_extend2(q, t, int(h0), wu, eb, opt=opt)
            cells += c
            if res[0] == prev or res[5] < (wu >> 1) + (wu >> 2):
                break
            prev = res[0]
        out.append(res + (wu, cells))
    return out


def check(ctx, pairs, h0s, ws, eb, flags=0, scoring=None):
    from quasimodo_b200 import _lib
    from quasimodo_b200.api import pack_ext_tasks
    from oracle import qmo_py
    og, oo = _lib.default_opt(), qmo_py.default_opt()
    for k, v in (scoring or {}).items():
        setattr(og, k, v)
        setattr(oo, k, v)
    seq, tasks = pack_ext_tasks(pairs, h0s, ws, eb, flags)
    got = ctx.extend_batch_host(seq, tasks, opt=og)
    want = oracle_results(pairs, h0s, ws, eb, retry=bool(flags & 1), prev_h0=bool(flags & 2), opt=oo)
    bad = []
    for i, wnt in enumerate(want):
        g = tuple(int(got[i][f]) for f in FIELDS) + (int(got[i]["w_used"]), int(got[i]["cells"]))
        if g != wnt:
            bad.append((i, g, wnt, len(pairs[i][0]), len(pairs[i][1]), int(h0s[i]), int(ws[i])))
    assert not bad, f"{len(bad)} of {len(want)} tasks differ; first: {bad[:3]}"


@pytest.mark.parametrize("eb", [0, 5])
def test_extend_random(ctx, eb):
    rng = np.random.default_rng(1234 + eb)
    pairs, h0s, ws = extgen.random_tasks(rng, 4000)
    check(ctx, pairs, h0s, ws, eb)


@pytest.mark.parametrize("scoring", [dict(a=2, b=5), dict(o_del=5, e_del=2, o_ins=7, e_ins=1), dict(a=3, b=4, o_del=4, e_del=3, o_ins=9, e_ins=2),
                                     dict(a=1, b=1, o_del=1, e_del=1, o_ins=1, e_ins=1)])
@pytest.mark.parametrize("n", [5000, 40000])
def test_extend_other_scoring_schemes(ctx, scoring, n):
    """match scores above 1 and unequal gap costs; 5,000 tasks run on the warp-per-task kernel, 40,000 (more than 4,096 per
    query-length class) on the thread-per-task kernel.  (A match score of 2 used to break the "dead diagonal" shortcut
    h + min(s, h) of both kernels: a cell with H = 1 on its diagonal got 1 + 1 instead of 1 + 2.)"""
    rng = np.random.default_rng(5)
    pairs, h0s, ws = extgen.random_tasks(rng, n)
    h0s = rng.integers(1, 250, len(h0s))
    check(ctx, pairs, h0s, ws, 5, scoring=scoring)


def test_extend_adversarial(ctx):
    pairs, h0s, ws = extgen.adversarial_tasks()
    check(ctx, pairs, h0s, ws, 5)
    check(ctx, pairs, h0s, ws, 0)


def test_extend_long_queries(ctx):
    rng = np.random.default_rng(77)
    pairs, h0s, ws = extgen.random_tasks(rng, 600, max_qlen=500)
    check(ctx, pairs, h0s, ws, 5)


@pytest.mark.parametrize("flags", [1, 3])
def test_extend_band_retry(ctx, flags):
    rng = np.random.default_rng(4321)
    pairs, h0s, ws = extgen.random_tasks(rng, 2000)
    ws = np.where(ws > 50, 10, ws)        # small bands so that the 2w retry actually triggers
    check(ctx, pairs, h0s, ws, 5, flags=flags)


def test_extend_rejects_bad_task(ctx):
    from quasimodo_b200 import QmError
    from quasimodo_b200.api import pack_ext_tasks
    seq, tasks = pack_ext_tasks([(np.zeros(600, np.uint8), np.zeros(10, np.uint8))], [31], [100], 5)
    with pytest.raises(QmError):
        ctx.extend_batch_host(seq, tasks)


def test_scalar_kernel_variant_is_bit_exact():
    """the default path is the packed two-tasks-per-thread kernel (extend3.cu); QM_EXT3=0 keeps the scalar thread-per-task
    kernel (extend2.cu: scores above 255, exotic scoring schemes) -- the same tasks through it; the switch is read once per
    process, hence the subprocess"""
    import os
    import subprocess
    import sys
    if os.environ.get("QM_EXT3"):
        pytest.skip("already inside the QM_EXT3 run")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "pytest", "tests/test_extend_gpu.py", "-q", "-x", "-k", "random or adversarial or band_retry or other_scoring"],
                       cwd=root, env=dict(os.environ, QM_EXT3="0"), capture_output=True, text=True)
    assert p.returncode == 0 and " passed" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_packed_kernel_equals_its_host_build(ctx):
    """ext3_kernel on tasks sorted by query length as the pipeline hands them over (the host build of the same statements is
    checked against the oracle in tests/test_ext3_host.py), scores up to and beyond the packed kernel's limit of 255: the
    tasks beyond it come back right through the fallback list"""
    rng = np.random.default_rng(2024)
    pairs, h0s, ws = extgen.random_tasks(rng, 60000, max_qlen=125)
    h0s = rng.integers(1, 250, len(h0s))
    order = np.argsort([len(q) for q, _ in pairs], kind="stable")
    pairs = [pairs[i] for i in order]
    h0s, ws = h0s[order], ws[order]
    check(ctx, pairs, h0s, ws, 5, flags=3)
