// ext3_host.cpp -- HOST build of the packed two-tasks-per-thread extension logic (quasimodo_b200/csrc/ext3_core.cuh) for the
// CPU test suite: the statements the CUDA kernel runs, with the DPX instructions emulated, driven task pair by task pair.
// Test infrastructure (built by tests/test_ext3_host.py into tests/_build/); the product has no host path.
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#define E3_STATS 1
#include "../quasimodo_b200/csrc/ext3_core.cuh"
E3Stats g_e3_stats = {0, 0, 0, 0, 0, 0, 0};
extern "C" void ext3_host_stats(long long *out, int reset) { out[0] = g_e3_stats.rows; out[1] = g_e3_stats.core_cols; out[2] = g_e3_stats.masked_cols; out[3] = g_e3_stats.solo_rows; if (reset) g_e3_stats = E3Stats{0, 0, 0, 0, 0, 0, 0}; }

namespace {
struct WideMem {                        // 16-bit planes: {h2, e2} 64-bit word + 32-bit query word per column (SmemWide of extend3.cu)
    static constexpr bool kNarrow = false;
    struct Raw { unsigned h2, e2, q2; };
    std::vector<Raw> v;
    explicit WideMem(int cap) : v((size_t)cap + 1 + kE3Pad) {}
    Raw raw(int j) { return v.at((size_t)j); }
    static void unpack(const Raw &r, unsigned &h2, unsigned &e2, unsigned &q2) { h2 = r.h2; e2 = r.e2; q2 = r.q2; }
    void put(int j, unsigned h2, unsigned e2) { v.at((size_t)j).h2 = h2; v.at((size_t)j).e2 = e2; }
    void set_he(int j, int X, int h, int e) { ((uint16_t *)&v.at((size_t)j).h2)[X] = (uint16_t)h; ((uint16_t *)&v.at((size_t)j).e2)[X] = (uint16_t)e; }
    bool zero(int j, int X) { return (((uint16_t *)&v.at((size_t)j).h2)[X] | ((uint16_t *)&v.at((size_t)j).e2)[X]) == 0; }
    void set_q(int j, int X, unsigned code) { ((uint16_t *)&v.at((size_t)j).q2)[X] = (uint16_t)code; }
};
struct NarrowMem {                      // byte planes: one 32-bit word (h | e << 8 per task) + a 16-bit query word per column (SmemNarrow)
    static constexpr bool kNarrow = true;
    struct Raw { unsigned he; unsigned q; };
    std::vector<unsigned> ehv;
    std::vector<uint16_t> qv;
    explicit NarrowMem(int cap) : ehv((size_t)cap + 1 + kE3Pad), qv((size_t)cap + kE3Pad) {}
    Raw raw(int j) { Raw r; r.he = ehv.at((size_t)j); r.q = (size_t)j < qv.size() ? qv[(size_t)j] : 0; return r; }
    static void unpack(const Raw &r, unsigned &h2, unsigned &e2, unsigned &q2)
    {
        h2 = r.he & 0x00ff00ffu; e2 = (r.he >> 8) & 0x00ff00ffu; q2 = (r.q & 0xffu) | (r.q & 0xff00u) << 8;
    }
    void put(int j, unsigned h2, unsigned e2) { ehv.at((size_t)j) = (h2 & 0x00ff00ffu) | (e2 & 0x00ff00ffu) << 8; }
    void set_he(int j, int X, int h, int e) { ((uint16_t *)&ehv.at((size_t)j))[X] = (uint16_t)((h & 0xff) | (e & 0xff) << 8); }
    bool zero(int j, int X) { return ((uint16_t *)&ehv.at((size_t)j))[X] == 0; }
    void set_q(int j, int X, unsigned code) { ((uint8_t *)&qv.at((size_t)j))[X] = (uint8_t)code; }
};
struct HostTgt {
    const uint8_t *t[2];
    int n[2];
    int raw(int X, int i) const { if (i < 0 || i >= n[X]) __builtin_trap(); return t[X][i]; }
    int decode(int, int c) const { return c > 4 ? 4 : c; }
};
struct HostQry { const uint8_t *q; int code(int j) const { return q[j] > 4 ? 4 : q[j]; } };
}  // namespace

// tasks i: query seq + q_off[i] (qlen[i]), target seq + t_off[i] (tlen[i]); flags bit 0 = band retry, bit 1 = prev starts at h0.
// Tasks are run in pairs (2k, 2k+1) in the given order; dirty != 0 leaves the previous pair's cells in place (what a
// persistent kernel thread sees), otherwise the planes start zeroed.  out[i] = 8 ints; not_ok[i] = 1: the packed kernel
// refuses the task (the launcher's fallback), out untouched.  Returns 0, or -1 when the scoring scheme is refused.
template <class HostMem>
static int host_run(const int *scores /* a b o_del e_del o_ins e_ins zdrop */, int cap, int64_t n, const uint8_t *seq,
                    const int64_t *q_off, const int64_t *t_off, const int *qlen, const int *tlen, const int *h0, const int *w,
                    int end_bonus, const unsigned *flags, int *out, uint8_t *not_ok)
{
    E3Scores S = {scores[0], scores[1], scores[2], scores[3], scores[4], scores[5], scores[6]};
    if (HostMem::kNarrow ? !e3_scores_ok_narrow(S) : !e3_scores_ok(S)) return -1;
    const E3Consts K = e3_consts(S);
    const bool sym = S.o_del == S.o_ins && S.e_del == S.e_ins && S.a == 1;
    HostMem mem(cap);
    for (int64_t i0 = 0; i0 < n; i0 += 2) {
        E3Half H[2];
        HostTgt tgt;
        for (int X = 0; X < 2; ++X) {
            H[X].tk = -1; H[X].phase = 0; tgt.t[X] = nullptr; tgt.n[X] = 0;
            const int64_t i = i0 + X;
            if (i >= n) continue;
            not_ok[i] = !e3_task_ok(S, qlen[i], h0[i], cap);
            if (not_ok[i]) continue;
            H[X].tk = (int)i; H[X].phase = 1;
            H[X].qlen = qlen[i]; H[X].tlen = tlen[i]; H[X].h0 = h0[i]; H[X].w0 = w[i]; H[X].w = w[i]; H[X].end_bonus = end_bonus;
            H[X].tries_left = (flags[i] & 1u) ? 2 : 1;
            H[X].prev = (flags[i] & 2u) ? h0[i] : -1;
            H[X].cells = 0;
            tgt.t[X] = seq + t_off[i]; tgt.n[X] = tlen[i];
            HostQry qry = {seq + q_off[i]};
            e3_load_query(K, qlen[i], X, mem, qry);
        }
        while (H[0].phase || H[1].phase) {
            for (int X = 0; X < 2; ++X) if (H[X].phase == 1) e3_start_try(K, H[X], X, mem, tgt);
            bool dA = false, dB = false;
            if (sym) e3_row<true>(K, H[0], H[1], mem, tgt, dA, dB);
            else e3_row<false>(K, H[0], H[1], mem, tgt, dA, dB);
            const bool d[2] = {dA, dB};
            for (int X = 0; X < 2; ++X)
                if (d[X]) {
                    E3Result r;
                    if (e3_end_try(H[X], &r)) { memcpy(out + 8 * (int64_t)H[X].tk, &r, sizeof r); H[X].tk = -1; }
                }
        }
    }
    return 0;
}

// narrow != 0: the byte-plane layout (needs a + b <= 16, returns -1 otherwise), else the 16-bit planes
extern "C" int ext3_host_run(const int *scores, int cap, int64_t n, const uint8_t *seq, const int64_t *q_off, const int64_t *t_off,
                             const int *qlen, const int *tlen, const int *h0, const int *w, int end_bonus, const unsigned *flags, int *out,
                             uint8_t *not_ok, int narrow)
{
    return narrow ? host_run<NarrowMem>(scores, cap, n, seq, q_off, t_off, qlen, tlen, h0, w, end_bonus, flags, out, not_ok)
                  : host_run<WideMem>(scores, cap, n, seq, q_off, t_off, qlen, tlen, h0, w, end_bonus, flags, out, not_ok);
}

// Warp model (analysis tool for the kernel's launch shape, not a test of results): `lanes` threads walk the task list the way
// ext3_kernel does (a thread takes two consecutive tasks; idle threads refill once `refill` of them are idle or nobody works)
// and step their rows in lockstep.  out[0] = row steps, out[1] = sum over steps of max-over-lanes common columns,
// out[2] = sum over steps of max-over-lanes one-task-only columns, out[3] = useful pair columns (common + one-task-only, all
// lanes), out[4] = steps in which some lane had one-task-only columns, out[5] = sum over steps of active lanes.
extern "C" int ext3_host_warpsim(const int *scores, int cap, int64_t n, const uint8_t *seq, const int64_t *q_off, const int64_t *t_off,
                                 const int *qlen, const int *tlen, const int *h0, const int *w, int end_bonus, const unsigned *flags,
                                 int lanes, int refill, long long *out)
{
    E3Scores S = {scores[0], scores[1], scores[2], scores[3], scores[4], scores[5], scores[6]};
    if (!e3_scores_ok(S)) return -1;
    const E3Consts K = e3_consts(S);
    const bool sym = S.o_del == S.o_ins && S.e_del == S.e_ins && S.a == 1;
    using HostMem = WideMem;
    struct Lane { HostMem mem; E3Half H[2]; HostTgt tgt; explicit Lane(int c) : mem(c) { H[0].phase = H[1].phase = 0; H[0].tk = H[1].tk = -1; } };
    std::vector<Lane> L;
    for (int l = 0; l < lanes; ++l) L.emplace_back(cap);
    int64_t cursor = 0;
    for (int k = 0; k < 6; ++k) out[k] = 0;
    for (;;) {
        int n_idle = 0, n_want = 0;
        for (auto &x : L) { const bool idle = !x.H[0].phase && !x.H[1].phase; n_idle += idle; n_want += idle && cursor < n; }
        if (n_idle == lanes && cursor >= n) break;
        if (n_want >= refill || n_idle == lanes)
            for (auto &x : L) {
                if (x.H[0].phase || x.H[1].phase || cursor >= n) continue;
                for (int X = 0; X < 2 && cursor < n; ++X, ++cursor) {
                    const int64_t i = cursor;
                    if (!e3_task_ok(S, qlen[i], h0[i], cap)) continue;
                    E3Half &H = x.H[X];
                    H.tk = (int)i; H.phase = 1; H.qlen = qlen[i]; H.tlen = tlen[i]; H.h0 = h0[i]; H.w0 = w[i]; H.w = w[i]; H.end_bonus = end_bonus;
                    H.tries_left = (flags[i] & 1u) ? 2 : 1; H.prev = (flags[i] & 2u) ? h0[i] : -1; H.cells = 0;
                    x.tgt.t[X] = seq + t_off[i]; x.tgt.n[X] = tlen[i];
                    HostQry qry = {seq + q_off[i]};
                    e3_load_query(K, qlen[i], X, x.mem, qry);
                }
            }
        int mx_core = 0, mx_masked = 0, active = 0;
        for (auto &x : L) {
            for (int X = 0; X < 2; ++X) if (x.H[X].phase == 1) e3_start_try(K, x.H[X], X, x.mem, x.tgt);
            if (!x.H[0].phase && !x.H[1].phase) continue;
            ++active;
            g_e3_stats.last_pre = g_e3_stats.last_core = 0;
            bool dA = false, dB = false;
            if (sym) e3_row<true>(K, x.H[0], x.H[1], x.mem, x.tgt, dA, dB);
            else e3_row<false>(K, x.H[0], x.H[1], x.mem, x.tgt, dA, dB);
            mx_core = std::max(mx_core, g_e3_stats.last_core); mx_masked = std::max(mx_masked, g_e3_stats.last_pre);
            out[3] += g_e3_stats.last_core + g_e3_stats.last_pre;
            E3Result r;
            if (dA && e3_end_try(x.H[0], &r)) x.H[0].tk = -1;
            if (dB && e3_end_try(x.H[1], &r)) x.H[1].tk = -1;
        }
        if (active) { out[0] += 1; out[1] += mx_core; out[2] += mx_masked; out[4] += mx_masked > 0; out[5] += active; }
    }
    return 0;
}
