"""Independent pure-Python restatement of ksw_extend2, used ONLY to
differential-test oracle/qmo_ksw.c (two restatements written separately from the same spec, as
SURVEY.md section 7 'hard part 6' asks).  Test infrastructure; small cases only."""


def score(a, b, t, q):
    if t > 3 or q > 3:
        return -1
    return a if t == q else -b


def ksw_extend2(query, target, h0, w, end_bonus, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100):
    qlen, tlen = len(query), len(target)
    H = [0] * (qlen + 1)
    E = [0] * (qlen + 1)
    H[0] = h0
    if qlen >= 1:
        H[1] = h0 - (o_ins + e_ins) if h0 > o_ins + e_ins else 0
    j = 2
    while j <= qlen and H[j - 1] > e_ins:
        H[j] = H[j - 1] - e_ins
        j += 1
    maxsc = max(a, -b, -1)
    w = min(w, max(1, int((qlen * maxsc + end_bonus - o_ins) / e_ins + 1.0)),
            max(1, int((qlen * maxsc + end_bonus - o_del) / e_del + 1.0)))
    mx, mx_i, mx_j, mx_ie, gscore, max_off = h0, -1, -1, -1, -1, 0
    beg, end = 0, qlen
    cells = 0
    for i in range(tlen):
        f, m, mj = 0, 0, -1
        beg = max(beg, i - w)
        end = min(end, i + w + 1, qlen)
        h1 = max(0, h0 - (o_del + e_del * (i + 1))) if beg == 0 else 0
        j = beg
        while j < end:
            M, e = H[j], E[j]
            H[j] = h1
            M = M + score(a, b, target[i], query[j]) if M else 0
            h = max(M, e, f)
            h1 = h
            if not (m > h):
                mj = j
            m = max(m, h)
            e = max(e - e_del, max(M - (o_del + e_del), 0))
            E[j] = e
            f = max(f - e_ins, max(M - (o_ins + e_ins), 0))
            j += 1
        cells += max(0, end - beg)
        H[end] = h1
        E[end] = 0
        if j == qlen:
            if not (gscore > h1):
                mx_ie = i
            gscore = max(gscore, h1)
        if m == 0:
            break
        if m > mx:
            mx, mx_i, mx_j = m, i, mj
            max_off = max(max_off, abs(mj - i))
        elif zdrop > 0:
            if i - mx_i > mj - mx_j:
                if mx - m - ((i - mx_i) - (mj - mx_j)) * e_del > zdrop:
                    break
            else:
                if mx - m - ((mj - mx_j) - (i - mx_i)) * e_ins > zdrop:
                    break
        j = beg
        while j < end and H[j] == 0 and E[j] == 0:
            j += 1
        beg = j
        j = end
        while j >= beg and H[j] == 0 and E[j] == 0:
            j -= 1
        end = min(j + 2, qlen)
    return (mx, mx_j + 1, mx_i + 1, mx_ie + 1, gscore, max_off), cells


def _local_matrix(query, target, a, b, o_del, e_del, o_ins, e_ins):
    """Full H matrix of the affine local alignment (rows = target, columns = query), Gotoh form."""
    n, m = len(target), len(query)
    NEG = -10 ** 9
    H = [[0] * (m + 1) for _ in range(n + 1)]
    E = [[NEG] * (m + 1) for _ in range(n + 1)]     # gap consuming target only (deletion), comes from the row above
    F = [[NEG] * (m + 1) for _ in range(n + 1)]     # gap consuming query only (insertion), comes from the left
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            E[i][j] = max(E[i - 1][j] - e_del, H[i - 1][j] - o_del - e_del)
            F[i][j] = max(F[i][j - 1] - e_ins, H[i][j - 1] - o_ins - e_ins)
            H[i][j] = max(0, H[i - 1][j - 1] + score(a, b, target[i - 1], query[j - 1]), E[i][j], F[i][j])
    return H


def ksw_align2(query, target, minsc, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1):
    """Independent restatement of ksw_align2 with KSW_XSUBO | KSW_XSTART (see oracle/qmo_ksw.c): works on the full
    matrix instead of rolling rows.  Returns (score, te, qe, score2, te2, tb, qb)."""
    def scan(q, t, minsc_, endsc):
        H = _local_matrix(q, t, a, b, o_del, e_del, o_ins, e_ins)
        best, te, qe, log = 0, -1, -1, []
        for i in range(len(t)):
            row = H[i + 1][1:]
            imax = max(row) if row else 0
            if imax >= minsc_:
                if not log or log[-1][1] + 1 != i:
                    log.append([imax, i])
                elif log[-1][0] < imax:
                    log[-1] = [imax, i]
            if imax > best:
                best, te, qe = imax, i, row.index(imax)
                if best >= endsc:
                    break
        return best, te, qe, log
    sc, te, qe, log = scan(query, target, minsc if minsc >= 0 else 1 << 16, 1 << 16)
    score2, te2 = -1, -1
    mx = (sc + max(a, 1) - 1) // max(a, 1)
    for s, r in log:
        if (r < te - mx or r > te + mx) and s > score2:
            score2, te2 = s, r
    tb = qb = -1
    if minsc >= 0 and sc >= minsc and te >= 0:
        s2, t2, q2, _ = scan(query[:qe + 1][::-1], target[:te + 1][::-1], 1 << 16, sc)
        if s2 == sc:
            tb, qb = te - t2, qe - q2
    return sc, te, qe, score2, te2, tb, qb


NEG_INF = -0x40000000


def ksw_global2(query, target, w, a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1):
    """Banded global alignment with traceback, written as WHOLE MATRICES and a traceback that
    re-derives every decision from the matrix values -- oracle/qmo_ksw.c keeps one rolling row and a byte of decision bits per
    cell.  M[i][j] = "arrive diagonally" (bwa's m), E[i][j] / F[i][j] = the best score ending in a deletion / insertion when
    cell (i, j) is entered (gaps open from m, not from h); cells outside the band are minus infinity.
    -> (score, [(op, len)...]) with op 0 = M, 1 = I, 2 = D"""
    qlen, tlen = len(query), len(target)
    oe_del, oe_ins = o_del + e_del, o_ins + e_ins

    def band(i):
        return max(0, i - w), min(qlen, i + w + 1)

    def h_above(i, j):                  # H[i][j] for i = -1 or j = -1 (the borders), minus infinity outside the band
        if i == -1 and j == -1:
            return 0
        if i == -1:
            return -(o_ins + e_ins * (j + 1)) if j + 1 <= w else NEG_INF
        return -(o_del + e_del * (i + 1)) if band(i)[0] == 0 else NEG_INF

    H, M, E, F = {}, {}, {}, {}
    for i in range(tlen):
        lo, hi = band(i)
        for j in range(lo, hi):
            if i > 0 and j > 0:
                diag = H.get((i - 1, j - 1), NEG_INF)
            elif i == 0:
                diag = h_above(-1, j - 1)
            else:
                diag = h_above(i - 1, -1)
            M[i, j] = diag + score(a, b, target[i], query[j])
            # deletion: extend the one that ended at (i-1, j), or open from the diagonal arrival there
            E[i, j] = max(E[i - 1, j] - e_del, M[i - 1, j] - oe_del) if (i - 1, j) in M else NEG_INF
            # insertion: likewise along the row; the first cell of a row's band has none
            F[i, j] = max(F[i, j - 1] - e_ins, M[i, j - 1] - oe_ins) if j > lo else NEG_INF
            H[i, j] = max(M[i, j], E[i, j], F[i, j])
    if tlen == 0:
        final = h_above(-1, qlen - 1)
    elif band(tlen - 1)[1] == qlen and qlen > 0:
        final = H[tlen - 1, qlen - 1]
    elif qlen == 0:
        final = h_above(tlen - 1, -1)
    else:
        final = NEG_INF                  # the band never reaches the last column (the rolling row's cell holds minus infinity)
    ops = []

    def push(op, n=1):
        if ops and ops[-1][0] == op:
            ops[-1][1] += n
        else:
            ops.append([op, n])

    i = tlen - 1
    k = min(i + w + 1, qlen) - 1
    state = 0
    while i >= 0 and k >= 0:
        m, e, f = M[i, k], E[i, k], F[i, k]
        if state == 0:                  # how was H[i][k] reached: diagonal wins ties over deletion, both over insertion
            state = 0 if m >= e else 1
            if max(m, e) < f:
                state = 2
        elif state == 1:                # we came up a deletion column: was E[i+1][k] an extension of E[i][k] or opened from m?
            state = 1 if e - e_del > m - oe_del else 0
        else:
            state = 2 if f - e_ins > m - oe_ins else 0
        if state == 0:
            push(0)
            i -= 1
            k -= 1
        elif state == 1:
            push(2)
            i -= 1
        else:
            push(1)
            k -= 1
    if i >= 0:
        push(2, i + 1)
    if k >= 0:
        push(1, k + 1)
    return final, [(op, n) for op, n in reversed(ops)]
