// pileup.cu -- per-column allele counting with the read admission and mate-overlap rules of
// `bcftools mpileup -B` (reference call site rules/vcfcall.smk:115; upstream htslib sam.c
// bam_plp_push / overlap_push / tweak_overlap_quality and bcftools bam2bcf.c bcf_call_glfgen;
// semantics SURVEY.md A.8-A.9; depth cap disabled, see DESIGN.md "deviations").
//
// One warp per read pair: the mate-overlap quality rewrite is pair-local, so both mates' qualities sit
// in shared memory (BAM SEQ orientation), the lanes walk the query bases, and every counted base is one
// fire-and-forget `red.global.add.s32` into the channel-major count tensor counts[ch][l_pac].  With
// consecutive lanes on consecutive reference positions of one channel plane the 32 reductions of a warp
// fall into one or two 128-byte lines of the L2-resident tensor.
#include "pipeline.cuh"

namespace {

constexpr int kMaxLen = 512;          // read length limit of this kernel (BASELINE reads are 150 / 250)
constexpr int kWarpsPerBlock = 4;

struct AlnS {                         // the fields of a qm_aln the walkers need, in shared memory
    int32_t pos, n_cigar;
    int32_t m0, m1;                   // gap-free alignment ([clip] M [clip], the common case): query bases [m0, m1) are the
                                      // M bases, base i sits at pos + i - m0; m1 = 0: not gap-free, walk the CIGAR
    uint32_t cigar[QM_MAX_CIGAR];
};

// [clip] <len>M [clip] ?  -> the query interval of the M bases
__device__ __forceinline__ void simple_span(const uint32_t *cigar, int nc, int32_t *m0, int32_t *m1)
{
    int k = 0, x = 0;
    *m0 = 0; *m1 = 0;
    if (k < nc && (cigar[k] & 0xf) == 4) { x = (int)(cigar[k] >> 4); ++k; }
    if (k >= nc || (cigar[k] & 0xf) != 0) return;
    const int len = (int)(cigar[k] >> 4);
    ++k;
    if (k < nc && (cigar[k] & 0xf) == 4) ++k;
    if (k != nc) return;
    *m0 = x; *m1 = x + len;
}

// reference position (contig coordinate) of query base i in SEQ order, or -1 when i is not an M-type base
__device__ __forceinline__ int rpos_of(const AlnS &a, int i)
{
    if (a.m1) return i >= a.m0 && i < a.m1 ? a.pos + i - a.m0 : -1;
    int x = 0, p = a.pos;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, len = (int)(a.cigar[k] >> 4);
        if (op == 0) { if (i < x + len) return p + (i - x); x += len; p += len; }
        else if (op == 1 || op == 4) { if (i < x + len) return -1; x += len; }
        else if (op == 2) p += len;
    }
    return -1;
}
// query index (SEQ order) of the M-type base aligned to contig position p, or -1
__device__ __forceinline__ int qidx_of(const AlnS &a, int p)
{
    if (a.m1) return p >= a.pos && p < a.pos + (a.m1 - a.m0) ? a.m0 + (p - a.pos) : -1;
    int x = 0, pp = a.pos;
    for (int k = 0; k < a.n_cigar; ++k) {
        const int op = a.cigar[k] & 0xf, len = (int)(a.cigar[k] >> 4);
        if (op == 0) { if (p >= pp && p < pp + len) return x + (p - pp); x += len; pp += len; }
        else if (op == 1 || op == 4) x += len;
        else if (op == 2) { if (p < pp + len) return -1; pp += len; }
    }
    return -1;
}

__device__ __forceinline__ void red_add(int32_t *p) { atomicAdd(p, 1); }   // result unused => RED.ADD

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pileup_kernel(IndexView V, qm_pileup_opt po, const qm_aln *__restrict__ alns, const uint8_t *__restrict__ codes,
              const uint8_t *__restrict__ quals, int stride, const int32_t *__restrict__ lens, int64_t n_pairs,
              int32_t *__restrict__ counts, unsigned long long *__restrict__ n_admitted)
{
    __shared__ AlnS s_aln[kWarpsPerBlock][2];
    __shared__ uint8_t s_q[kWarpsPerBlock][2][kMaxLen];
    const int lane = qm_lane(), wib = threadIdx.x >> 5;
    const int64_t warp0 = blockIdx.x * (int64_t)kWarpsPerBlock + wib;
    const int64_t n_warps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int64_t L_pac = V.l_pac;
    unsigned long long admitted_local = 0;

    for (int64_t pi = warp0; pi < n_pairs; pi += n_warps) {
        const qm_aln *g[2] = { alns + 2 * pi, alns + 2 * pi + 1 };
        int flag[2], rid[2], tlen[2], L[2];
        bool ok[2], rev[2];
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint32_t fm = *(const uint32_t *)&g[e]->flag;           // flag | mapq<<16 | n_cigar<<24
            flag[e] = fm & 0xffff;
            const int mapq = (fm >> 16) & 0xff, nc = fm >> 24;
            rid[e] = g[e]->rid; tlen[e] = g[e]->tlen; L[e] = lens[2 * pi + e];
            rev[e] = (flag[e] & 0x10) != 0;
            ok[e] = !(flag[e] & (0x4 | 0x100 | 0x200 | 0x400)) && nc != 0 && nc != 255 && mapq >= po.min_mapq &&
                    !((flag[e] & 0x1) && !(flag[e] & 0x2) && !po.count_orphans) && L[e] <= kMaxLen;
            if (ok[e]) {
                if (lane == 0) {
                    s_aln[wib][e].pos = g[e]->pos; s_aln[wib][e].n_cigar = nc;
                    simple_span(g[e]->cigar, nc, &s_aln[wib][e].m0, &s_aln[wib][e].m1);
                }
                if (lane < nc) s_aln[wib][e].cigar[lane] = g[e]->cigar[lane];
                const uint8_t *qv = quals + (2 * pi + e) * stride;
                for (int i = lane; i < L[e]; i += 32) s_q[wib][e][i] = rev[e] ? qv[L[e] - 1 - i] : qv[i];
            }
        }
        __syncwarp();
        const uint8_t *rd[2] = { codes + (2 * pi) * stride, codes + (2 * pi + 1) * stride };
        auto seq_base = [&](int e, int i) {
            const int c = rev[e] ? rd[e][L[e] - 1 - i] : rd[e][i];
            return rev[e] ? (c > 3 ? 4 : 3 - c) : c;
        };
        // ---- mate overlap: the mate that comes first in coordinate order plays htslib's `a` ----
        if (!po.ignore_overlaps && ok[0] && ok[1] && rid[0] == rid[1] && (flag[0] & 0x2) && !(flag[0] & 0x8) &&
            abs(tlen[0]) < 2 * L[0] && abs(tlen[1]) < 2 * L[1]) {
            const int p0 = s_aln[wib][0].pos, p1 = s_aln[wib][1].pos;
            const int A = (p1 < p0 || (p1 == p0 && (int)rev[1] < (int)rev[0])) ? 1 : 0, B = A ^ 1;
            for (int ia = lane; ia < L[A]; ia += 32) {
                const int p = rpos_of(s_aln[wib][A], ia);
                if (p < 0) continue;
                const int ib = qidx_of(s_aln[wib][B], p);
                if (ib < 0) continue;
                const int qa = s_q[wib][A][ia], qb = s_q[wib][B][ib];
                if (seq_base(A, ia) == seq_base(B, ib)) {
                    const int q = qa + qb;
                    s_q[wib][A][ia] = (uint8_t)(q > 200 ? 200 : q); s_q[wib][B][ib] = 0;
                } else if (qa >= qb) { s_q[wib][A][ia] = (uint8_t)(0.8 * qa); s_q[wib][B][ib] = 0; }
                else { s_q[wib][B][ib] = (uint8_t)(0.8 * qb); s_q[wib][A][ia] = 0; }
            }
        }
        __syncwarp();
        // ---- counting ----
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            if (!ok[e]) continue;
            ++admitted_local;
            const AlnS &a = s_aln[wib][e];
            const int64_t base = V.off[rid[e]];
            for (int i = lane; i < L[e]; i += 32) {
                const int p = rpos_of(a, i);
                if (p < 0) continue;
                red_add(counts + 14 * L_pac + base + p);
                if (s_q[wib][e][i] >= po.min_bq) red_add(counts + (int64_t)((rev[e] ? 6 : 0) + seq_base(e, i)) * L_pac + base + p);
            }
            if (lane < a.n_cigar) {          // operation-level channels: lane k owns CIGAR operation k
                int p = a.pos, last_m = -1;
                bool started = false;
                for (int k = 0; k < lane; ++k) {
                    const int op = a.cigar[k] & 0xf, len = (int)(a.cigar[k] >> 4);
                    if (op == 0) { p += len; last_m = p - 1; started = true; }
                    else if (op == 2) p += len;
                }
                const int op = a.cigar[lane] & 0xf, len = (int)(a.cigar[lane] >> 4);
                if (op == 0 && !started) red_add(counts + 15 * L_pac + base + p);
                else if (op == 1) { if (last_m >= 0) red_add(counts + 12 * L_pac + base + last_m); }
                else if (op == 2) {
                    if (last_m >= 0) red_add(counts + 13 * L_pac + base + last_m);
                    for (int i = 0; i < len; ++i) red_add(counts + (int64_t)(rev[e] ? 11 : 5) * L_pac + base + p + i);
                }
            }
        }
    }
    if (n_admitted && lane == 0 && admitted_local) atomicAdd(n_admitted, admitted_local);
}

// [16][l_pac] planes -> [l_pac][16] rows (the count-TSV row order)
__global__ void planes_to_rows_kernel(const int32_t *__restrict__ planes, int64_t l_pac, int32_t *__restrict__ rows)
{
    __shared__ int32_t tile[QM_NCH][33];
    const int64_t p0 = blockIdx.x * 32ll;
    for (int c = threadIdx.y; c < QM_NCH; c += blockDim.y) {
        const int64_t p = p0 + threadIdx.x;
        tile[c][threadIdx.x] = p < l_pac ? planes[c * l_pac + p] : 0;
    }
    __syncthreads();
    const int t = threadIdx.y * 32 + threadIdx.x;          // 256 threads: 16 positions x 16 channels per pass
    for (int r = t; r < 32 * QM_NCH; r += 256) {
        const int pos = r / QM_NCH, c = r % QM_NCH;
        if (p0 + pos < l_pac) rows[(p0 + pos) * QM_NCH + c] = tile[c][pos];
    }
}

}  // namespace

extern "C" {

void qm_pileup_opt_default(qm_pileup_opt *p)
{
    p->min_mapq = 0; p->min_bq = 13; p->count_orphans = 0; p->ignore_overlaps = 0;
}

int qm_pileup_accumulate(qm_ctx *ctx, const qm_index *idx, const qm_pileup_opt *po, const qm_aln *d_alns,
                         const uint8_t *d_codes, const uint8_t *d_quals, int32_t stride, const int32_t *d_lens,
                         int64_t n_pairs, int32_t *d_counts, void *stream)
{
    if (!ctx || !idx || !po || n_pairs < 0 || (n_pairs > 0 && (!d_alns || !d_codes || !d_quals || !d_lens || !d_counts)))
        return QM_EINVAL;
    if (n_pairs == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t blocks = (n_pairs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)ctx->sm_count * 16;       // persistent: 16 blocks x 4 warps per SM
    if (blocks > cap) blocks = cap;
    const int sp = qm_prof_begin(ctx, QM_ST_PILEUP, (cudaStream_t)stream);
    pileup_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        idx->v, *po, d_alns, d_codes, d_quals, stride, d_lens, n_pairs, d_counts, nullptr);
    qm_prof_end(ctx, QM_ST_PILEUP, sp, (cudaStream_t)stream, 1);
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

int qm_counts_to_rows(qm_ctx *ctx, const qm_index *idx, const int32_t *d_planes, int32_t *d_rows, void *stream)
{
    if (!ctx || !idx || !d_planes || !d_rows) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t l_pac = idx->v.l_pac;
    planes_to_rows_kernel<<<(unsigned)((l_pac + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(d_planes, l_pac, d_rows);
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

}  // extern "C"
