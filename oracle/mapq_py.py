"""ORACLE (test infrastructure) -- a second restatement of how bwa-mem finishes a read whose pair cannot be judged (no usable
insert-size model): bwamem.c mem_mark_primary_se (the order of the hits, which one is primary, its sub-optimal score and the
number of near-equal rivals) and mem_approx_mapq_se (the mapping quality), per SURVEY.md A.5.  Written to be diffed against
oracle/qmo_mem.c on batches too small for a model (< 10 pairs), where every record is finished single-end."""
import math

M64 = (1 << 64) - 1


def hash_64(key):
    """Thomas Wang's 64-bit mix, as bwa uses it to break score ties reproducibly"""
    key = (key + (~(key << 32) & M64)) & M64
    key ^= key >> 22
    key = (key + (~(key << 13) & M64)) & M64
    key ^= key >> 8
    key = (key + (key << 3)) & M64
    key ^= key >> 15
    key = (key + (~(key << 27) & M64)) & M64
    key ^= key >> 31
    return key


def finish_single_end(regs, n, read_id, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, min_seed_len=31, T=30, mask_level=0.5,
                      coef_len=50):
    """regs[:n] = the read's hits as the single-end stage left them; read_id = pair index * 2 + mate.
    -> None (unmapped) or dict(score, sub (the XS value), mapq, rb) of the primary record"""
    if n == 0:
        return None
    hits = [dict(score=int(r["score"]), qb=int(r["qb"]), qe=int(r["qe"]), rb=int(r["rb"]), re=int(r["re"]), csub=int(r["csub"]),
                 tie=hash_64((read_id + i) & M64), sub=0, rivals=0) for i, r in enumerate(regs[:n])]
    hits.sort(key=lambda h: (-h["score"], h["tie"]))
    near = max(a + b, o_del + e_del, o_ins + e_ins)
    heads = [hits[0]]                                      # hits that no better hit overlaps on the read
    for h in hits[1:]:
        for p in heads:
            lo, hi = max(p["qb"], h["qb"]), min(p["qe"], h["qe"])
            if hi > lo and hi - lo >= min(h["qe"] - h["qb"], p["qe"] - p["qb"]) * mask_level:
                if p["sub"] == 0:
                    p["sub"] = h["score"]
                if p["score"] - h["score"] <= near:
                    p["rivals"] += 1
                break
        else:
            heads.append(h)
    best = hits[0]
    if best["score"] < T:
        return None
    sub = best["sub"] if best["sub"] else min_seed_len * a
    sub = max(sub, best["csub"])
    if sub >= best["score"]:
        mapq = 0
    else:
        span = max(best["qe"] - best["qb"], best["re"] - best["rb"])
        identity = 1. - (span * a - best["score"]) / (a + b) / span
        scale = 1. if span < coef_len else math.log(coef_len) / math.log(span)
        scale *= identity * identity
        mapq = int(6.02 * (best["score"] - sub) / a * scale * scale + .499)
        if best["rivals"] > 0:
            mapq -= int(4.343 * math.log(best["rivals"] + 1) + .499)
        mapq = min(60, max(0, mapq))
    return dict(score=best["score"], sub=max(best["sub"], best["csub"]), mapq=mapq, rb=best["rb"])
