// baq_core.cuh -- base alignment quality, the per-read logic: htslib 1.9 realn.c sam_prob_realn (window, agreement with the CIGAR,
// extended BAQ, the cap) over probaln.c kpa_glocal (banded profile HMM in double precision).  Plain C++ over a scratch accessor,
// compiled for the device by baq.cu and for the HOST by tests/baq_host.cpp: the CPU test suite runs the very statements the kernels
// run against the oracle (oracle/qmo_baq.c), byte for byte, without a GPU.  See baq.cu for the layout on the GPU and the notes on
// bit-exactness (no FMA contraction, IEEE division, the float error table, log()).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <math.h>
#include "../../include/quasimodo_b200.h"

#if defined(__CUDACC__)
#define BAQ_HD __host__ __device__
#else
#define BAQ_HD
#endif

constexpr double kBaqEI = .25, kBaqEM = .33333333333;

struct BaqRead { int32_t read, xb, xe, bw; };

// the reads the pileup counts (pileup.cu): BAQ of any other read is never looked at
BAQ_HD inline bool baq_admitted(const qm_pileup_opt &po, const qm_aln &a)
{
    if (a.flag & (0x4 | 0x100 | 0x200 | 0x400)) return false;
    if (a.n_cigar == 0 || a.n_cigar == 255) return false;
    if ((int)a.mapq < po.min_mapq) return false;
    if ((a.flag & 0x1) && !(a.flag & 0x2) && !po.count_orphans) return false;
    return true;
}

// the reference window and band of one read (sam_prob_realn's first half); false = the read is left alone
BAQ_HD inline bool baq_window(const qm_aln &a, int l_qseq, int64_t ref_len, int &xb_o, int &xe_o, int &bw_o)
{
    int x = a.pos, y = 0, yb = -1, ye = -1, xb = -1, xe = -1;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, l = (int)(a.cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (yb < 0) yb = y;
            if (xb < 0) xb = x;
            ye = y + l; xe = x + l;
            x += l; y += l;
        } else if (op == 4 || op == 1) y += l;
        else if (op == 2) x += l;
        else if (op == 3) return false;
    }
    if (xb < 0 || l_qseq <= 0) return false;
    int bw = 7;
    const int dd = abs((xe - xb) - (ye - yb));
    if (dd > bw) bw = dd + 3;
    xb -= yb + bw / 2; if (xb < 0) xb = 0;
    xe += l_qseq - ye + bw / 2;
    if (xe - xb - l_qseq > bw) {
        xb += (xe - xb - l_qseq - bw) / 2;
        xe -= (xe - xb - l_qseq - bw) / 2;               // sees the line above's xb, as upstream
    }
    if (xe > ref_len) xe = (int)ref_len;
    if (xe - xb <= 0) return false;
    xb_o = xb; xe_o = xe; bw_o = bw;
    return true;
}

// the band the HMM really uses (kpa_glocal's first lines)
BAQ_HD inline int baq_band(int l_ref, int l_query, int cbw)
{
    int bw = l_ref > l_query ? l_ref : l_query;
    if (bw > cbw) bw = cbw;
    if (bw < abs(l_ref - l_query)) bw = abs(l_ref - l_query);
    return bw;
}

#define SET_U(u, b, i, k) { int x_ = (i) - (b); x_ = x_ > 0 ? x_ : 0; (u) = ((k) - x_ + 1) * 3; }

// probaln.c kpa_glocal on the scratch of accessor A.  ref(k), qry(i), err(i): 1-based reference code, query code, error
// probability (float) of the read in BAM orientation.  out_state / out_q (i = 0 .. l_query-1) receive the decoding.
template <class Acc, class Ref, class Qry, class Err, class Out>
BAQ_HD void baq_glocal(const Acc &A, int l_ref, int l_query, int cbw, double cd, double ce, Ref ref, Qry qry, Err err, Out out)
{
    const int bw = baq_band(l_ref, l_query, cbw);
    const int bw2 = bw * 2 + 1, row = bw2 * 3 + 6;
    double m[9];
    const double sM = 1. / (2 * l_query + 2), sI = sM;
    m[0] = (1 - cd - cd) * (1 - sM); m[1] = m[2] = cd * (1 - sM);
    m[3] = (1 - ce) * (1 - sI); m[4] = ce * (1 - sI); m[5] = 0.;
    m[6] = 1 - ce; m[7] = 0.; m[8] = ce;
    const double bM = (1 - cd) / l_ref, bI = cd / l_ref;
    // every cell the recurrences may look at starts as zero (upstream: calloc)
    for (int i = 0; i <= l_query; ++i) for (int c = 0; c < row; ++c) A.f(i, c) = 0.;
    for (int j = 0; j < 2; ++j) for (int c = 0; c < row; ++c) A.b(j, c) = 0.;
    int k;
    /*** forward ***/
    SET_U(k, bw, 0, 0);
    A.f(0, k) = 1.; A.s(0) = 1.;
    {
        double sum = 0.;
        const int beg = 1, end = l_ref < bw + 1 ? l_ref : bw + 1;
        const int q1 = qry(1);
        const double ql = (double)err(1);
        for (k = beg; k <= end; ++k) {
            int u;
            const int rk = ref(k);
            const double e = (rk > 3 || q1 > 3) ? 1. : rk == q1 ? 1. - ql : ql * kBaqEM;
            SET_U(u, bw, 1, k);
            const double f0 = e * bM, f1 = kBaqEI * bI;
            A.f(1, u) = f0; A.f(1, u + 1) = f1;
            sum += f0 + f1;
        }
        A.s(1) = sum;
        int _beg, _end;
        SET_U(_beg, bw, 1, beg); SET_U(_end, bw, 1, end); _end += 2;
        for (k = _beg; k <= _end; ++k) A.f(1, k) = A.f(1, k) / sum;
    }
    for (int i = 2; i <= l_query; ++i) {
        double sum = 0.;
        const double qli = (double)err(i);
        int beg = 1, end = l_ref, x, _beg, _end;
        const int qyi = qry(i);
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = beg; k <= end; ++k) {
            int u, v11, v01, v10;
            const int rk = ref(k);
            const double e = (rk > 3 || qyi > 3) ? 1. : rk == qyi ? 1. - qli : qli * kBaqEM;
            SET_U(u, bw, i, k); SET_U(v11, bw, i - 1, k - 1); SET_U(v10, bw, i - 1, k); SET_U(v01, bw, i, k - 1);
            const double f0 = e * (m[0] * A.f(i - 1, v11) + m[3] * A.f(i - 1, v11 + 1) + m[6] * A.f(i - 1, v11 + 2));
            const double f1 = kBaqEI * (m[1] * A.f(i - 1, v10) + m[4] * A.f(i - 1, v10 + 1));
            const double f2 = m[2] * A.f(i, v01) + m[8] * A.f(i, v01 + 2);
            A.f(i, u) = f0; A.f(i, u + 1) = f1; A.f(i, u + 2) = f2;
            sum += f0 + f1 + f2;
        }
        A.s(i) = sum;
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        sum = 1. / sum;
        for (k = _beg; k <= _end; ++k) A.f(i, k) = A.f(i, k) * sum;
    }
    {
        double sum = 0.;
        for (k = 1; k <= l_ref; ++k) {
            int u;
            SET_U(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            sum += A.f(l_query, u) * sM + A.f(l_query, u + 1) * sI;
        }
        A.s(l_query + 1) = sum;
    }
    /*** backward + posterior decoding, row by row: row i lives in b(i & 1) ***/
    auto decode = [&](int i) {
        double sum = 0., mx = 0.;
        int beg = 1, end = l_ref, x, max_k = -1;
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (int kk = beg; kk <= end; ++kk) {
            int u;
            SET_U(u, bw, i, kk);
            double z = A.f(i, u) * A.b(i & 1, u); if (z > mx) { mx = z; max_k = (kk - 1) << 2 | 0; } sum += z;
            z = A.f(i, u + 1) * A.b(i & 1, u + 1); if (z > mx) { mx = z; max_k = (kk - 1) << 2 | 1; } sum += z;
        }
        mx /= sum;
        int qq = (int)(-4.343 * log(1. - mx) + .499);
        out(i - 1, max_k, qq > 100 ? 99 : qq);
    };
    {
        const int j = l_query & 1;
        const double sl = A.s(l_query), sl1 = A.s(l_query + 1);
        for (k = 1; k <= l_ref; ++k) {
            int u;
            SET_U(u, bw, l_query, k);
            if (u < 3 || u >= bw2 * 3 + 3) continue;
            A.b(j, u) = sM / sl / sl1; A.b(j, u + 1) = sI / sl / sl1;
        }
        decode(l_query);
    }
    for (int i = l_query - 1; i >= 1; --i) {
        const int j = i & 1, j1 = j ^ 1;
        for (int c = 0; c < row; ++c) A.b(j, c) = 0.;                   // this slot held row i + 2
        int beg = 1, end = l_ref, x, _beg, _end;
        double y = (i > 1);
        const double qli1 = (double)err(i + 1);
        const int qyi1 = qry(i + 1);
        x = i - bw; beg = beg > x ? beg : x;
        x = i + bw; end = end < x ? end : x;
        for (k = end; k >= beg; --k) {
            int u, v11, v01, v10;
            SET_U(u, bw, i, k); SET_U(v11, bw, i + 1, k + 1); SET_U(v10, bw, i + 1, k); SET_U(v01, bw, i, k + 1);
            double e;
            if (k >= l_ref) e = 0;
            else { const int rk = ref(k + 1); e = (rk > 3 || qyi1 > 3) ? 1. : rk == qyi1 ? 1. - qli1 : qli1 * kBaqEM; }
            e = e * A.b(j1, v11);
            const double b10 = A.b(j1, v10 + 1), b01 = A.b(j, v01 + 2);
            A.b(j, u) = e * m[0] + kBaqEI * m[1] * b10 + m[2] * b01;
            A.b(j, u + 1) = e * m[3] + kBaqEI * m[4] * b10;
            A.b(j, u + 2) = (e * m[6] + m[8] * b01) * y;
        }
        SET_U(_beg, bw, i, beg); SET_U(_end, bw, i, end); _end += 2;
        y = 1. / A.s(i);
        for (k = _beg; k <= _end; ++k) A.b(j, k) = A.b(j, k) * y;
        decode(i);
    }
}

// One read: window, HMM, then sam_prob_realn's second half (agreement with the CIGAR, extended BAQ inside each M block, the
// cap).  st / qv: per-read scratch for the decoding (state int32, phred byte), strided by `ss`.
template <class Acc>
BAQ_HD void baq_one(const uint8_t *refw /* the contig's bases from xb on */, const float *q2p /* 10^(-q/10), 256 floats */, const Acc &A,
                    const BaqRead br, const qm_aln &a, const uint8_t *codes, const uint8_t *quals, int stride, int L, int flag,
                    uint8_t *quals_out, int32_t *st, uint8_t *qv, size_t ss)
{
    const bool rev = (a.flag & 0x10) != 0;
    const uint8_t *rd = codes + (size_t)br.read * stride, *ql = quals + (size_t)br.read * stride;
    uint8_t *qo = quals_out + (size_t)br.read * stride;
    auto ref = [&](int k) { const int c = refw[k - 1]; return c > 3 ? 4 : c; };
    auto qry = [&](int i) { const int c = rev ? rd[L - i] : rd[i - 1]; return c > 3 ? 4 : (rev ? 3 - c : c); };
    auto err = [&](int i) { return q2p[rev ? ql[L - i] : ql[i - 1]]; };
    auto out = [&](int i, int state, int q) { st[(size_t)i * ss] = state; qv[(size_t)i * ss] = (uint8_t)q; };
    baq_glocal(A, br.xe - br.xb, L, br.bw, 0.001, 0.1, ref, qry, err, out);
    const bool extend = (flag >> 1) & 1;
    // BAM-orientation base i is the read's base (rev ? L-1-i : i); the output keeps the read's own orientation
    auto qual_at = [&](int i) { return (int)(rev ? ql[L - 1 - i] : ql[i]); };
    auto put = [&](int i, int bq) {
        const int q0 = qual_at(i);
        // qual -= (64 + (qual <= bq ? 0 : qual - bq)) - 64   (extended)   |   qual -= (qual - bq + 64) - 64   (plain; bq <= qual)
        qo[rev ? L - 1 - i : i] = (uint8_t)(extend ? (q0 <= bq ? q0 : bq) : bq);
    };
    int x = a.pos, y = 0;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, l = (int)(a.cigar[k] >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (!extend) {
                for (int i = y; i < y + l; ++i) {
                    const int s = st[(size_t)i * ss], q = qv[(size_t)i * ss], q0 = qual_at(i);
                    const int bq = ((s & 3) != 0 || s >> 2 != x - br.xb + (i - y)) ? 0 : (q0 < q ? q0 : q);
                    put(i, bq);
                }
            } else {
                // bq = agreement ? q : 0; left = running max from the block's start (kept in qv), right = from its end
                int run = 0;
                for (int i = y; i < y + l; ++i) {
                    const int s = st[(size_t)i * ss], q = qv[(size_t)i * ss];
                    const int bq = ((s & 3) != 0 || s >> 2 != x - br.xb + (i - y)) ? 0 : q;
                    st[(size_t)i * ss] = bq;
                    run = i == y ? bq : (bq > run ? bq : run);
                    qv[(size_t)i * ss] = (uint8_t)run;
                }
                for (int i = y + l - 1; i >= y; --i) {
                    const int bq = st[(size_t)i * ss];
                    run = i == y + l - 1 ? bq : (bq > run ? bq : run);
                    const int left = qv[(size_t)i * ss];
                    put(i, left < run ? left : run);
                }
            }
            x += l; y += l;
        } else if (op == 4 || op == 1) y += l;
        else if (op == 2) x += l;
    }
}

