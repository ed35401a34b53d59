// depthcap.cu -- the per-file depth cap of `bcftools mpileup -d N` (rules/vcfcall.smk:115 runs with the default; SURVEY.md
// A.8, 8f-2): which admitted reads htslib's pileup iterator keeps.  htslib 1.9 sam.c: bam_plp_push refuses a read whose start
// equals the position the iterator is assembling while more than `maxcnt` nodes are in use (the buffered reads plus the
// iterator's spare tail node); bam_plp_next releases the reads that end at or before that position, then steps to the next
// position or jumps to the first buffered read; bam_plp_auto feeds one read whenever nothing lies beyond the position.
// The outcome depends on the order of the reads by construction, so this is a replay of the iterator over the records in BAM
// order: keys and record ends come from the device (sorted there, sort.cu), the replay itself is a host loop of O(reads +
// positions) steps -- a serial dependence from read to read that has no parallel form -- and the drop mask goes back to the
// device for the counting kernel.  Off by default (at BASELINE depths the cap throws away most of the sample).
#include <queue>
#include <vector>
#include "pipeline.cuh"

namespace {

__device__ __forceinline__ bool cap_admitted(const qm_pileup_opt &po, const qm_aln &a)
{
    const int nc = a.n_cigar;
    return !(a.flag & (0x4 | 0x100 | 0x200 | 0x400)) && nc != 0 && nc != 255 && a.mapq >= po.min_mapq &&
           !((a.flag & 0x1) && !(a.flag & 0x2) && !po.count_orphans);
}

__global__ void __launch_bounds__(256)
cap_info_kernel(qm_pileup_opt po, const qm_aln *__restrict__ alns, int64_t n, int64_t r0, uint64_t *__restrict__ keys, int32_t *__restrict__ ends)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const qm_aln a = alns[i];
    uint64_t key = ~0ull >> 16;                         // not admitted: sorts behind every record (48 key bits)
    int end = 0;
    if (cap_admitted(po, a)) {
        int rlen = 0;
        for (int k = 0; k < a.n_cigar; ++k) { const int op = a.cigar[k] & 0xf; if (op == 0 || op == 2) rlen += (int)(a.cigar[k] >> 4); }
        end = a.pos + (rlen > 0 ? rlen : 1);
        key = (uint64_t)(uint32_t)a.rid << 33 | (uint64_t)(uint32_t)a.pos << 1 | (uint64_t)((a.flag & 0x10) != 0);
    }
    keys[r0 + i] = key;
    ends[r0 + i] = end;
}

}  // namespace

extern "C" {

// chunks: the sample's records in input order.  d_drop[c]: 2 * n_pairs[c] bytes, 1 = the iterator dropped the read.  Synchronous.
int qm_depth_cap(qm_ctx *ctx, const qm_pileup_opt *po, int n_chunks, const qm_aln *const *d_alns, const int64_t *n_pairs, int max_depth,
                 uint8_t *const *d_drop, int64_t *h_n_dropped, void *stream)
{
    if (!ctx || !po || n_chunks < 0 || max_depth < 1 || (n_chunks > 0 && (!d_alns || !n_pairs || !d_drop))) return QM_EINVAL;
    QM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t N = 0;
    for (int c = 0; c < n_chunks; ++c) { if (n_pairs[c] < 0) return QM_EINVAL; N += 2 * n_pairs[c]; }
    if (h_n_dropped) *h_n_dropped = 0;
    if (N == 0) return QM_OK;
    if (N > 0x7fffffffll) return qm_fail(ctx, QM_ELIMIT, "qm_depth_cap: more than 2^31 records");
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_keys = 0, o_keys0 = al((size_t)N * 8), o_ends = o_keys0 + al((size_t)N * 8), o_perm = o_ends + al((size_t)N * 4);
    void *p = nullptr;
    int rc = qm_scratch_reserve(ctx, 24, o_perm + al((size_t)N * 4), &p);
    if (rc) return rc;
    char *b = (char *)p;
    uint64_t *keys = (uint64_t *)(b + o_keys), *keys0 = (uint64_t *)(b + o_keys0);
    int32_t *ends = (int32_t *)(b + o_ends);
    uint32_t *perm = (uint32_t *)(b + o_perm);
    int64_t r0 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t n = 2 * n_pairs[c];
        if (n) cap_info_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(*po, d_alns[c], n, r0, keys, ends);
        r0 += n;
    }
    QM_CUDA(ctx, cudaGetLastError());
    QM_CUDA(ctx, cudaMemcpyAsync(keys0, keys, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    rc = qm_sort_pairs(ctx, keys, perm, N, 48, st);                    // stable: BAM order, ties in input order
    if (rc) return rc;
    std::vector<uint64_t> h_keys((size_t)N);
    std::vector<uint32_t> h_perm((size_t)N);
    std::vector<int32_t> h_ends((size_t)N);
    QM_CUDA(ctx, cudaMemcpyAsync(h_keys.data(), keys, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaMemcpyAsync(h_perm.data(), perm, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaMemcpyAsync(h_ends.data(), ends, (size_t)N * 4, cudaMemcpyDeviceToHost, st));
    QM_CUDA(ctx, cudaStreamSynchronize(st));

    // ---- the replay ----
    std::vector<uint8_t> drop((size_t)N, 0);
    struct Node { int tid; int64_t beg, end; bool gone; };
    std::vector<Node> nodes;                                           // buffered reads in arrival order
    using HeapItem = std::pair<std::pair<int, int64_t>, size_t>;       // ((contig, end), node)
    std::priority_queue<HeapItem, std::vector<HeapItem>, std::greater<HeapItem>> by_end;
    size_t head = 0;
    int cur_tid = 0, last_tid = -1;
    int64_t cur_pos = 0, last_pos = -1, in_use = 1;                    // the spare tail node counts
    int64_t n_dropped = 0;
    const uint64_t not_admitted = ~0ull >> 16;
    for (int64_t i = 0; i < N; ++i) {
        const uint64_t key = h_keys[(size_t)i];
        if (key == not_admitted) break;                                // everything behind is not admitted either
        const size_t rec = h_perm[(size_t)i];
        const int tid = (int)(key >> 33);
        const int64_t beg = (int64_t)((key >> 1) & 0xffffffffull), end = h_ends[rec];
        while (last_tid > cur_tid || (last_tid == cur_tid && last_pos > cur_pos)) {
            while (!by_end.empty() && (by_end.top().first.first < cur_tid || (by_end.top().first.first == cur_tid && by_end.top().first.second <= cur_pos))) {
                nodes[by_end.top().second].gone = true;
                --in_use;
                by_end.pop();
            }
            while (head < nodes.size() && nodes[head].gone) ++head;
            if (head < nodes.size() && cur_tid < nodes[head].tid) { cur_tid = nodes[head].tid; cur_pos = nodes[head].beg; }
            else if (head < nodes.size() && cur_pos < nodes[head].beg) cur_pos = nodes[head].beg;
            else ++cur_pos;
        }
        if (cur_tid == tid && cur_pos == beg && in_use > max_depth) { drop[rec] = 1; ++n_dropped; continue; }
        last_tid = tid; last_pos = beg;
        if (end > cur_pos || tid > cur_tid) {
            by_end.push({{tid, end}, nodes.size()});
            nodes.push_back({tid, beg, end, false});
            ++in_use;
        }
    }
    r0 = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t n = 2 * n_pairs[c];
        if (n) QM_CUDA(ctx, cudaMemcpyAsync(d_drop[c], drop.data() + r0, (size_t)n, cudaMemcpyHostToDevice, st));
        r0 += n;
    }
    QM_CUDA(ctx, cudaStreamSynchronize(st));
    if (h_n_dropped) *h_n_dropped = n_dropped;
    return QM_OK;
}

}  // extern "C"
