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


def order_and_mark(regs, n, read_id, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, mask_level=0.5):
    """mem_mark_primary_se: the read's hits by (score descending, bwa's hash of read id + arrival index); a hit overlapped on the read
    by a better one is secondary to it (`parent` = its index), the better one notes the first such score (`sub`) and counts the
    rivals within one edit's worth of score"""
    hits = [dict(score=int(r["score"]), qb=int(r["qb"]), qe=int(r["qe"]), rb=int(r["rb"]), re=int(r["re"]), csub=int(r["csub"]),
                 rid=int(r["rid"]), truesc=int(r["truesc"]), w=int(r["w"]), tie=hash_64((read_id + i) & M64), sub=0, rivals=0, parent=-1) for i, r in enumerate(regs[:n])]
    hits.sort(key=lambda h: (-h["score"], h["tie"]))
    near = max(a + b, o_del + e_del, o_ins + e_ins)
    heads = [0] if hits else []
    for i in range(1, len(hits)):
        h = hits[i]
        for j in heads:
            p = hits[j]
            lo, hi = max(p["qb"], h["qb"]), min(p["qe"], h["qe"])
            if hi > lo and hi - lo >= min(h["qe"] - h["qb"], p["qe"] - p["qb"]) * mask_level:
                if p["sub"] == 0:
                    p["sub"] = h["score"]
                if p["score"] - h["score"] <= near:
                    p["rivals"] += 1
                h["parent"] = j
                break
        else:
            heads.append(i)
    return hits


def approx_mapq(h, a=1, b=4, min_seed_len=31, coef_len=50):
    """mem_approx_mapq_se (frac_rep = 0: the hash seeder has no repetitive-seed fraction)"""
    sub = max(h["sub"] if h["sub"] else min_seed_len * a, h["csub"])
    if sub >= h["score"]:
        return 0
    span = max(h["qe"] - h["qb"], h["re"] - h["rb"])
    identity = 1. - (span * a - h["score"]) / (a + b) / span
    scale = 1. if span < coef_len else math.log(coef_len) / math.log(span)
    scale *= identity * identity
    mapq = int(6.02 * (h["score"] - sub) / a * scale * scale + .499)
    if h["rivals"] > 0:
        mapq -= int(4.343 * math.log(h["rivals"] + 1) + .499)
    return min(60, max(0, mapq))


def finish_single_end(regs, n, read_id, T=30, **scoring):
    """regs[:n] = the read's hits as the single-end stage left them; read_id = pair index * 2 + mate.
    -> None (unmapped) or dict(score, sub (the XS value), mapq, rb) of the primary record"""
    hits = order_and_mark(regs, n, read_id, **scoring)
    if not hits or hits[0]["score"] < T:
        return None
    best = hits[0]
    return dict(score=best["score"], sub=max(best["sub"], best["csub"]), mapq=approx_mapq(best), rb=best["rb"], hit=best)
