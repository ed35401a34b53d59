"""ORACLE (test infrastructure) -- a second restatement of bwa-mem's pairing decision (bwamem_pair.c mem_pair and the MAPQ logic of
mem_sam_pe), to be diffed against oracle/qmo_mem.c with mate rescue off (the same hit lists go into both).

mem_pair is written here as what its sorted sweep computes: every pair of hits of the two mates, the earlier one (on the forward
strand's coordinate) first, whose orientation has a model and whose distance lies in the model's window, scored
score1 + score2 + 0.721 ln(2 erfc(|z| / sqrt 2)) a; the best pair wins, ties broken by bwa's hash."""
import math

from oracle.mapq_py import M64, approx_mapq, hash_64, order_and_mark


def _raw_mapq(diff, a):
    return int(6.02 * diff / a + .499)


def best_pair(l_pac, offs, pes, hits, pair_id, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1):
    """hits = [hits of mate 0, hits of mate 1] (ordered by order_and_mark) -> (score, second score, rivals, [index, index]) or None"""
    ends = []
    for m in (0, 1):
        for i, h in enumerate(hits[m]):
            rev = h["rb"] >= l_pac
            fwd = 2 * l_pac - 1 - h["rb"] if rev else h["rb"]
            ends.append(((h["rid"] << 32) | (fwd - int(offs[h["rid"]])), (h["score"] << 32) | (i << 2) | (int(rev) << 1) | m))
    ends.sort()
    cands = []
    for i in range(len(ends)):
        xi, yi = ends[i]
        for k in range(i):
            xk, yk = ends[k]
            if (yk & 1) == (yi & 1):
                continue                                   # two hits of the same mate
            d = ((yk >> 1 & 1) << 1) | (yi >> 1 & 1)       # orientation: strand of the earlier hit, strand of the later one
            if pes[d]["failed"]:
                continue
            dist = xi - xk
            if not (int(pes[d]["low"]) <= dist <= int(pes[d]["high"])):
                continue
            z = (dist - float(pes[d]["avg"])) / float(pes[d]["std"])
            q = max(0, int((yi >> 32) + (yk >> 32) + .721 * math.log(2. * math.erfc(abs(z) * math.sqrt(0.5))) * a + .499))
            tag = (k << 32) | i
            cands.append(((q << 32) | (hash_64((tag ^ (pair_id << 8)) & M64) & 0xffffffff), tag))
    if not cands:
        return None
    cands.sort()
    top = cands[-1]
    k, i = top[1] >> 32, top[1] & 0xffffffff
    chosen = [None, None]
    for e in (i, k):
        chosen[ends[e][1] & 1] = (ends[e][1] & 0xffffffff) >> 2
    second = cands[-2][0] >> 32 if len(cands) > 1 else 0
    near = max(a + b, o_del + e_del, o_ins + e_ins)
    rivals = sum(1 for c in cands[:-1] if second - (c[0] >> 32) <= near)
    return top[0] >> 32, second, rivals, chosen


def finish_pair(l_pac, offs, pes, regs0, n0, regs1, n1, pair_id, a=1, T=30, pen_unpaired=17):
    """-> (proper-pair flag, [None | dict(score, sub, mapq, rb) per mate])"""
    hits = [order_and_mark(regs0, n0, 2 * pair_id), order_and_mark(regs1, n1, 2 * pair_id + 1)]
    if hits[0] and hits[1]:
        found = best_pair(l_pac, offs, pes, hits, pair_id)
        ambiguous = any(any(h["parent"] < 0 and h["score"] >= T for h in hs[1:]) for hs in hits)
        if found and found[0] > 0 and not ambiguous:
            o, subo, n_sub, z = found
            unpaired = hits[0][0]["score"] + hits[1][0]["score"] - pen_unpaired
            subo = max(subo, unpaired)
            q_pe = _raw_mapq(o - subo, a)
            if n_sub > 0:
                q_pe -= int(4.343 * math.log(n_sub + 1) + .499)
            q_pe = min(60, max(0, q_pe))
            out = []
            if o > unpaired:
                for m in (0, 1):
                    c = hits[m][z[m]]
                    if c["parent"] >= 0:
                        c["sub"], c["parent"] = hits[m][c["parent"]]["score"], -2
                    q = approx_mapq(c)
                    q = q if q > q_pe else (q_pe if q_pe < q + 40 else q + 40)
                    q = min(q, _raw_mapq(c["score"] - c["csub"], a))
                    out.append(dict(score=c["score"], sub=max(c["sub"], c["csub"]), mapq=q, rb=c["rb"]))
                return True, out
            return False, [dict(score=h[0]["score"], sub=max(h[0]["sub"], h[0]["csub"]), mapq=approx_mapq(h[0]), rb=h[0]["rb"]) for h in hits]
    out = []
    for hs in hits:
        if hs and hs[0]["score"] >= T:
            out.append(dict(score=hs[0]["score"], sub=max(hs[0]["sub"], hs[0]["csub"]), mapq=approx_mapq(hs[0]), rb=hs[0]["rb"]))
        else:
            out.append(None)
    proper = False
    if out[0] and out[1] and hits[0][0]["rid"] == hits[1][0]["rid"]:
        b1, b2 = hits[0][0]["rb"], hits[1][0]["rb"]
        r1, r2 = b1 >= l_pac, b2 >= l_pac
        p2 = b2 if r1 == r2 else 2 * l_pac - 1 - b2
        d = (0 if r1 == r2 else 1) ^ (0 if p2 > b1 else 3)
        proper = (not pes[d]["failed"]) and int(pes[d]["low"]) <= abs(p2 - b1) <= int(pes[d]["high"])
    return proper, out
