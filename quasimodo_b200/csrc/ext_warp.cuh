// ext_warp.cuh -- ksw_extend2 executed by ONE WARP (formulation "B" of SURVEY.md A.3): one DP row per step, the
// query columns striped across the lanes in blocks of C consecutive columns, F as a max-plus warp scan.  Used by
// the low-latency kernels: ext_kernel<C> (extend.cu, small task lists) and tail_kernel (align.cu, the reads that are
// still active after the bulk rounds).  See extend.cu for the design notes.
#pragma once
#include "pipeline.cuh"

// where a task's bases come from: explicit byte strings (public C-ABI tasks) or the read batch /
// reference index in place (pipeline tasks; nothing is materialised in HBM)
struct SeqFetch {
    const uint8_t *q, *t;
    int64_t t0;
    int qstep, tstep;
    bool indirect;
    const IndexView *V;
    __device__ __forceinline__ int qbase(int j) const { return q[(int64_t)j * qstep]; }
    __device__ __forceinline__ int tbase(int i) const
    {
        return indirect ? qm_ref_base(*V, t0 + (int64_t)i * tstep) : t[i];
    }
};

struct ExtState {
    int score, qle, tle, gtle, gscore, max_off, cells;
};

// One ksw_extend2 call executed by one warp.  All lanes return the same ExtState.
template <int C>
__device__ __forceinline__ ExtState ext_run(const ExtParams &P, const SeqFetch &S, int qlen, int tlen, int h0, int w,
                                            int end_bonus, int lane)
{
    const unsigned FULL = 0xffffffffu;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int j0 = lane * C;

    int h[C], e[C], qc[C], mis[C];
#pragma unroll
    for (int k = 0; k < C; ++k) {
        const int j = j0 + k;
        // row -1 of eh[].h (SURVEY.md A.3 first three lines), closed form
        int v = 0;
        if (j == 0) v = h0;
        else if (j <= qlen) {
            const int vj = h0 - oe_ins - (j - 1) * P.e_ins;
            if (j == 1) v = vj > 0 ? vj : 0;
            else v = (vj + P.e_ins > P.e_ins) ? vj : 0;
        }
        h[k] = v;
        e[k] = 0;
        const int c = (j < qlen) ? S.qbase(j) : 4;
        qc[k] = c;
        mis[k] = (c > 3) ? -1 : -P.b;
    }

    {   // band cannot usefully exceed what the scores can pay for (doubles, as the reference)
        int best = P.a > -1 ? P.a : -1;
        if (-P.b > best) best = -P.b;
        int lim = (int)((double)(qlen * best + end_bonus - P.o_ins) / P.e_ins + 1.);
        lim = lim > 1 ? lim : 1;
        w = w < lim ? w : lim;
        lim = (int)((double)(qlen * best + end_bonus - P.o_del) / P.e_del + 1.);
        lim = lim > 1 ? lim : 1;
        w = w < lim ? w : lim;
    }

    int mx = h0, mx_i = -1, mx_j = -1, mx_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen, cells = 0;

    int tb_next = tlen > 0 ? S.tbase(0) : 0;
    for (int i = 0; i < tlen; ++i) {
        const int tb = tb_next;
        if (i + 1 < tlen) tb_next = S.tbase(i + 1);

        if (beg < i - w) {
            // columns that fall out of the band are never read again by the reference; zero them so
            // that they behave like the reference's trimmed (dead) columns.
            beg = i - w;
#pragma unroll
            for (int k = 0; k < C; ++k)
                if (j0 + k < beg) { h[k] = 0; e[k] = 0; }
        }
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1_init = 0;
        if (beg == 0) {
            h1_init = h0 - (P.o_del + P.e_del * (i + 1));
            h1_init = h1_init > 0 ? h1_init : 0;
        }
        const int na = end - j0;       // columns k < na of this lane are inside [.., end)
        const bool tN = tb > 3;

        // ---- pass 1: M, E(i+1,.), lane-local F chain with zero carry-in ----
        int M[C], fin[C], en[C];
        int f = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            int s = (qc[k] == tb) ? P.a : mis[k];
            if (tN) s = -1;
            const int hk = h[k];
            const int m = hk + min(s, hk << 8);     // == hk ? hk + s : <= 0  (a dead diagonal stays dead; scores are < 256)
            M[k] = m;
            en[k] = __viaddmax_s32_relu(e[k], -P.e_del, m - oe_del);
            fin[k] = f;
            f = __viaddmax_s32_relu(f, -P.e_ins, m - oe_ins);
        }
        // ---- warp max-plus scan of the F carry ----
        int carry = f;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(FULL, carry, d);
            carry = __viaddmax_s32(o, -d * C * P.e_ins, carry);   // lanes < d get their own value back: harmless
        }
        int Fin = __shfl_up_sync(FULL, carry, 1);
        if (lane == 0) Fin = 0;

        // ---- pass 2: H(i,.), row statistics ----
        int H[C];
        int key = 0;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            int hk = __vimax3_s32(M[k], e[k], fin[k]);
            hk = __viaddmax_s32(Fin, -k * P.e_ins, hk);
            hk = (k < na) ? hk : 0;
            H[k] = hk;
            key = max(key, (hk << 9) | (j0 + k));
        }
        // ---- commit eh[]: h shifts one column right, e in place; only indices <= end are written ----
        int hleft = __shfl_up_sync(FULL, H[C - 1], 1);
        if (lane == 0) hleft = h1_init;
        int nzlast = -1, nzfirst = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < C; ++k) {
            const int hl = (k == 0) ? hleft : H[k - 1];
            if (k <= na) h[k] = hl;
            e[k] = (k < na) ? en[k] : ((k == na) ? 0 : e[k]);
            const bool nz = (k <= na) && ((h[k] | e[k]) != 0);
            nzlast = nz ? (j0 + k) : nzlast;
            nzfirst = (nz && nzfirst == 0x7fffffff) ? (j0 + k) : nzfirst;
        }
        const int kmax = __reduce_max_sync(FULL, key);
        const int m = kmax >> 9, mj = kmax & 511;

        if (end > beg) cells += end - beg;
        const int jstop = end > beg ? end : beg;
        if (jstop == qlen) {
            int h1 = h1_init;
            if (end > beg) {
                int v = 0;
#pragma unroll
                for (int k = 0; k < C; ++k) v = (j0 + k == end - 1) ? H[k] : v;
                h1 = __reduce_max_sync(FULL, v);
            }
            mx_ie = gscore > h1 ? mx_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > mx) {
            mx = m; mx_i = i; mx_j = mj;
            const int d = abs(mj - i);
            max_off = max_off > d ? max_off : d;
        } else if (P.zdrop > 0) {
            const int dr = i - mx_i, dc = mj - mx_j;
            if (dr > dc) { if (mx - m - (dr - dc) * P.e_del > P.zdrop) break; }
            else         { if (mx - m - (dc - dr) * P.e_ins > P.zdrop) break; }
        }
        // ---- trim to the non-zero span of eh[beg..end] ----
        const int jl = __reduce_max_sync(FULL, nzlast);
        const int jf = __reduce_min_sync(FULL, nzfirst);
        beg = jf;
        end = jl + 2 < qlen ? jl + 2 : qlen;
    }
    ExtState r;
    r.score = mx; r.qle = mx_j + 1; r.tle = mx_i + 1; r.gtle = mx_ie + 1; r.gscore = gscore;
    r.max_off = max_off; r.cells = cells;
    return r;
}

