// simulate.cu -- C-ABI of the read simulator (host loop and device kernel share simulate.cuh).
#include "common.cuh"
#include "simulate.cuh"

namespace {
constexpr int kMaxSources = 8;
struct SimSources {
    int64_t off[kMaxSources], len[kMaxSources];
    uint32_t cum[kMaxSources];
};

__global__ void sim_kernel(qm_sim_params P, SimSources S, const uint8_t *__restrict__ genome, int64_t pair0,
                           int64_t n_pairs, int stride, uint8_t *__restrict__ codes, uint8_t *__restrict__ quals)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    uint8_t *b1 = codes + (2 * i) * stride, *b2 = b1 + stride;
    uint8_t *q1 = quals + (2 * i) * stride, *q2 = q1 + stride;
    qm_sim_pair(P, genome, S.off, S.len, S.cum, pair0 + i, b1, q1, b2, q2, nullptr, nullptr);
    for (int j = P.read_len; j < stride; ++j) { b1[j] = 4; b2[j] = 4; q1[j] = 0; q2[j] = 0; }
}

int check_params(const qm_sim_params *p, int stride)
{
    if (!p || p->read_len <= 0 || p->read_len > stride || p->n_sources <= 0 || p->n_sources > kMaxSources ||
        p->ins_max < p->read_len)
        return QM_EINVAL;
    return QM_OK;
}
}  // namespace

extern "C" {

int qm_simulate_pairs_host(const qm_sim_params *p, const uint8_t *h_genome, const int64_t *src_off,
                           const int64_t *src_len, const uint32_t *src_cum, int64_t pair0, int64_t n_pairs,
                           int32_t stride, uint8_t *h_codes, uint8_t *h_quals, int32_t *h_src, int64_t *h_pos)
{
    if (check_params(p, stride) || !h_genome || !src_off || !src_len || !src_cum || !h_codes || !h_quals || n_pairs < 0)
        return QM_EINVAL;
    for (int64_t i = 0; i < n_pairs; ++i) {
        uint8_t *b1 = h_codes + (2 * i) * stride, *b2 = b1 + stride;
        uint8_t *q1 = h_quals + (2 * i) * stride, *q2 = q1 + stride;
        qm_sim_pair(*p, h_genome, src_off, src_len, src_cum, pair0 + i, b1, q1, b2, q2,
                    h_src ? h_src + i : nullptr, h_pos ? h_pos + i : nullptr);
        for (int j = p->read_len; j < stride; ++j) { b1[j] = 4; b2[j] = 4; q1[j] = 0; q2[j] = 0; }
    }
    return QM_OK;
}

int qm_simulate_pairs(qm_ctx *ctx, const qm_sim_params *p, const uint8_t *d_genome, const int64_t *h_src_off,
                      const int64_t *h_src_len, const uint32_t *h_src_cum, int64_t pair0, int64_t n_pairs,
                      int32_t stride, uint8_t *d_codes, uint8_t *d_quals, void *stream)
{
    if (!ctx) return QM_EINVAL;
    if (check_params(p, stride) || !d_genome || !h_src_off || !h_src_len || !h_src_cum || !d_codes || !d_quals || n_pairs < 0)
        return qm_fail(ctx, QM_EINVAL, "qm_simulate_pairs: bad arguments");
    if (n_pairs == 0) return QM_OK;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    SimSources S;
    for (int s = 0; s < p->n_sources; ++s) { S.off[s] = h_src_off[s]; S.len[s] = h_src_len[s]; S.cum[s] = h_src_cum[s]; }
    const int tpb = 128;
    sim_kernel<<<(unsigned)((n_pairs + tpb - 1) / tpb), tpb, 0, (cudaStream_t)stream>>>(*p, S, d_genome, pair0, n_pairs, stride, d_codes, d_quals);
    QM_CUDA(ctx, cudaGetLastError());
    return QM_OK;
}

}  // extern "C"
