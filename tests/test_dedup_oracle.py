"""CPU sanity of the duplicate-marking restatement (oracle/dedup_py.py) on hand-made records."""
import numpy as np

from oracle import dedup_py, qmo_py


def rec(rid, pos, flag, cigar):
    a = np.zeros(1, dtype=qmo_py.ALN_DTYPE)[0]
    a["rid"], a["pos"], a["flag"], a["n_cigar"] = rid, pos, flag, len(cigar)
    for k, (op, ln) in enumerate(cigar):
        a["cigar"][k] = ln << 4 | "MIDNS".index(op)
    return a


def test_pairs_fragments_and_clips():
    M100 = [("M", 100)]
    pairs = [
        (rec(0, 1000, 0x63, M100), rec(0, 1200, 0x93, M100)),                       # 0: pair A
        (rec(0, 1000, 0x63, M100), rec(0, 1200, 0x93, M100)),                       # 1: same ends, better qualities -> stays
        (rec(0, 1003, 0x63, [("S", 3), ("M", 97)]), rec(0, 1200, 0x93, M100)),      # 2: clipped start, same unclipped 5' end -> duplicate
        (rec(0, 1000, 0x63, M100), rec(0, 1201, 0x93, M100)),                       # 3: other end differs by one -> not a duplicate
        (rec(0, 1000, 0x49, M100), rec(0, 1000, 0x85, [])),                          # 4: fragment at an end of a pair -> duplicate
        (rec(0, 5000, 0x49, M100), rec(0, 5000, 0x85, [])),                          # 5: fragment, alone with 6
        (rec(0, 5000, 0x49, M100), rec(0, 5000, 0x85, [])),                          # 6: fragment, same key, same score -> duplicate of 5
        (rec(-1, -1, 0x4d, []), rec(-1, -1, 0x8d, [])),                              # 7: unplaced pair
        (rec(0, 1200, 0x53, M100), rec(0, 1000, 0xa3, M100)),                       # 8: mates swapped, same ends as pair A -> duplicate
    ]
    alns = np.array([r for p in pairs for r in p], dtype=qmo_py.ALN_DTYPE)
    quals = np.full((len(alns), 100), 30, dtype=np.uint8)
    quals[2] = 40
    quals[3] = 40                                                                    # pair 1 scores highest
    lens = np.full(len(alns), 100, dtype=np.int32)
    dup = dedup_py.mark_duplicates(alns, quals, lens)
    assert dup.tolist() == [True, False, True, False, True, False, True, False, True]
    quals[:] = 30                                                                    # all equal: the first of the file stays
    dup = dedup_py.mark_duplicates(alns, quals, lens)
    assert dup.tolist() == [False, True, True, False, True, False, True, False, True]
    quals[0, :50] = 14                                                               # qualities below 15 do not count
    dup = dedup_py.mark_duplicates(alns, quals, lens)
    assert dup.tolist()[:3] == [True, False, True]


def test_dictionary_and_sorted_table_formulations_agree():
    """oracle/dedup_py.py (groups in dictionaries) against oracle/dedup_sort_py.py (one sorted table of read ends, runs of equal keys,
    as picard works) on a deep sample of a small genome -- thousands of positional duplicates, fragments whose mate is unplaced,
    soft-clipped ends -- and on the hand-made cases above"""
    from oracle import dedup_sort_py
    from quasimodo_b200 import workloads
    n = 12_000
    W = workloads.Workload("phix-deep", [("Phix", 1)], ["Phix"], n, 901)
    codes, quals, _, _ = W.simulate_host(0, n)
    lens = np.full(2 * n, 150, np.int32)
    codes[1:800:2] = 4                                       # 400 unplaced second mates
    codes[1001:1400:2, 120:] = (codes[1001:1400:2, 120:] + 1) % 4          # ruined tails: soft clips at 3' ends
    ref = qmo_py.Ref(W.ref.codes, W.ref.lens, k=31)
    alns = qmo_py.run_sample(ref, codes, quals, lens)[0]
    a, b = dedup_py.mark_duplicates(alns, quals, lens), dedup_sort_py.mark_duplicates(alns, quals, lens)
    assert np.array_equal(a, b), np.flatnonzero(a != b)[:10]
    clipped = sum(1 for x in alns if x["n_cigar"] not in (0, 255) and any((int(c) & 15) == 4 for c in x["cigar"][:x["n_cigar"]]))
    assert 0.02 < a.mean() < 0.9 and a[:400].any() and clipped > 50
