"""Shared task generators for the extension-kernel parity tests (random, adversarial, read-like)."""
import numpy as np


def mutate(rng, seq, sub, indel):
    out = []
    for b in seq:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(int(rng.integers(0, 4)))
        out.append(int(rng.integers(0, 4)) if rng.random() < sub else int(b))
    return np.array(out, dtype=np.uint8)


def random_tasks(rng, n, max_qlen=130):
    pairs, h0s, ws = [], [], []
    for _ in range(n):
        qlen = int(rng.integers(1, max_qlen))
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        kind = int(rng.integers(0, 5))
        if kind == 0:
            t = rng.integers(0, 4, int(rng.integers(0, 2 * qlen + 10))).astype(np.uint8)
        else:
            t = mutate(rng, q, [0.0, 0.02, 0.06, 0.15][kind - 1], [0.0, 0.005, 0.02, 0.06][kind - 1])
            t = np.concatenate([t, rng.integers(0, 4, int(rng.integers(0, qlen + 12))).astype(np.uint8)])
        if rng.random() < 0.1 and qlen > 2:
            q[rng.integers(0, qlen)] = 4
        if rng.random() < 0.05 and len(t) > 2:
            t[rng.integers(0, len(t))] = 4
        pairs.append((q, t))
        h0s.append(int(rng.integers(1, 160)))
        ws.append(int(rng.choice([1, 2, 3, 5, 20, 100, 200])))
    return pairs, np.array(h0s), np.array(ws)


def adversarial_tasks():
    """band-edge, z-drop boundary, all-N, zero-length, h0 extremes, long indels, w < qlen"""
    rng = np.random.default_rng(99)
    P, H, W = [], [], []

    def add(q, t, h0, w):
        P.append((np.array(q, dtype=np.uint8), np.array(t, dtype=np.uint8)))
        H.append(h0)
        W.append(w)

    base = rng.integers(0, 4, 120).astype(np.uint8)
    add([], [0, 1, 2], 31, 100)                               # empty query
    add([0, 1, 2], [], 31, 100)                               # empty target
    add([4] * 40, rng.integers(0, 4, 80), 50, 100)            # all-N query
    add(base[:40], [4] * 80, 50, 100)                         # all-N target
    add(base, base, 1, 100)                                   # minimal h0
    add(base, base, 150, 100)                                 # large h0
    add(base, np.concatenate([base[:60], base[75:]]), 60, 100)        # 15-base insertion in read
    add(base, np.concatenate([base[:60], rng.integers(0, 4, 30).astype(np.uint8), base[60:]]), 60, 100)  # 30-base deletion
    for w in (1, 2, 5, 10, 31, 32, 33):                       # w < qlen, band edges around lane boundaries
        add(base, np.concatenate([base, base[:20]]), 40, w)
        add(base[:33], np.concatenate([base[:33], base[:40]]), 120, w)
    # z-drop boundary: perfect prefix then garbage; lengths around the z-drop trigger
    for extra in range(95, 110):
        add(np.concatenate([base[:20], (3 - base[20:20 + extra]) % 4]), np.concatenate([base[:20], base[20:20 + extra]]), 31, 200)
    # h0 > w + 6 with qlen > w (SURVEY.md A.3 property 3: stale first-row values beyond the band)
    for w in (3, 8, 20):
        add(base, np.concatenate([base[:50], base[52:]]), w + 40, w)
        add(base, rng.integers(0, 4, 200), w + 100, w)
    # lengths at the striping class boundaries
    for ql in (31, 32, 63, 64, 127, 128, 255, 256, 300, 511):
        q = rng.integers(0, 4, ql).astype(np.uint8)
        add(q, mutate(rng, q, 0.04, 0.01), 45, 100)
        add(q, np.concatenate([q, q[:10]]), 31, 200)
    # long target after a short query (rows far beyond qlen + w)
    add(base[:10], rng.integers(0, 4, 400), 140, 100)
    add(base[:10], np.concatenate([base[:10], rng.integers(0, 4, 300).astype(np.uint8)]), 140, 5)
    return P, np.array(H), np.array(W)
