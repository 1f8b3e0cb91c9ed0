// Sequence-offset matcher: the body of the reference's evaluation hot loop
// eval/eval_faiss.py:204-232, batched over test ids.
//
//   seq_gather_kernel   q rows of every test id (id .. id+L-1; rows past the end repeat the last one, :208)
//   <segment search>    one top-k_probe search of all n_test*L rows (:211); rows of a shorter
//                       sequence length are a prefix of the longest one, so one search serves all
//   seq_cand_kernel     offset compensation (:215-216), sorted unique candidates >= 0 (:219) and,
//                       per candidate, the running mean of q[j].recon[c+j] (:222-229) for every
//                       requested length -- a gather/reduce over the fp32 rows
//   seq_top_kernel      the 10 best candidates per (test id, length), ties to the lower id (:232)
// Candidates this shard does not own score -inf, so that a row-sharded database can combine the
// per-rank score tables with a max-reduce between the two kernels.
#include <climits>

#include "index.h"
#include "ptx.cuh"

namespace nafp {

constexpr int SEQ_MAXC = 1024;     // k_probe * max_len upper bound
constexpr int SEQ_MAXL = 32;
constexpr int SEQ_NPRED = 10;

__global__ void seq_gather_kernel(const float* __restrict__ qall, int64_t n_query_rows,
                                  const int64_t* __restrict__ test_ids, int64_t n_test, int L,
                                  float* __restrict__ qrows) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_test * L) return;
    const int64_t t = w / L;
    const int j = static_cast<int>(w % L);
    // rows past the end of the query set (eval_faiss.py:208 truncates the slice) are never used by the
    // matcher; they repeat the last real row so that the segment search does not see degenerate
    // all-zero queries (every database row ties for those, which would force the exact fallback)
    int64_t src = test_ids[t] + j;
    if (src >= n_query_rows) src = n_query_rows - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (test_ids[t] >= 0 && test_ids[t] < n_query_rows) v = reinterpret_cast<const float4*>(qall + src * D128)[lane];
    reinterpret_cast<float4*>(qrows + w * D128)[lane] = v;
}

__device__ void bitonic_sort_asc_1024(uint64_t* keys) {
    for (int k = 2; k <= SEQ_MAXC; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < SEQ_MAXC; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint64_t a = keys[i], b = keys[ixj];
                    const bool asc = (i & k) == 0;
                    if (asc ? (a > b) : (a < b)) {
                        keys[i] = b;
                        keys[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(256)
seq_cand_kernel(const float* __restrict__ qrows, int64_t n_query_rows, const int64_t* __restrict__ test_ids,
                const int32_t* __restrict__ seq_lens, int n_len, int L, int k_probe,
                const int64_t* __restrict__ I, const float* __restrict__ x32, int64_t n_rows_local,
                int64_t n_rows_global, int64_t label_offset, int64_t owned_lo, int64_t owned_hi,
                int64_t* __restrict__ cand_ids, float* __restrict__ cand_scores, int32_t* __restrict__ n_cand) {
    __shared__ uint64_t keys[SEQ_MAXC];
    __shared__ int64_t uniq_c[SEQ_MAXC];
    __shared__ uint8_t uniq_j[SEQ_MAXC];
    __shared__ int wsum[8];
    __shared__ float prefix[8][SEQ_MAXL + 1];
    const int64_t t = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t id = test_ids[t];
    int lq = 0;
    if (id >= 0 && id < n_query_rows) lq = static_cast<int>(n_query_rows - id < L ? n_query_rows - id : L);

    for (int e = tid; e < SEQ_MAXC; e += blockDim.x) {
        uint64_t key = ~0ull;
        if (e < L * k_probe) {
            const int j = e / k_probe, m = e % k_probe;
            if (j < lq) {
                const int64_t lab = I[(t * L + j) * k_probe + m];
                const int64_t c = lab - j;
                if (lab >= 0 && c >= 0) key = (static_cast<uint64_t>(c) << 6) | static_cast<uint64_t>(j);
            }
        }
        keys[e] = key;
    }
    __syncthreads();
    bitonic_sort_asc_1024(keys);

    // compaction of first occurrences (sorted by candidate, then by j => first has the minimal j)
    int base = 0;
    for (int e0 = 0; e0 < SEQ_MAXC; e0 += blockDim.x) {
        const int e = e0 + tid;
        const uint64_t key = keys[e];
        const bool head = key != ~0ull && (e == 0 || (keys[e - 1] >> 6) != (key >> 6));
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += wsum[w];
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += wsum[w];
        if (head) {
            const int pos = off + __popc(bal & ((1u << lane) - 1));
            uniq_c[pos] = static_cast<int64_t>(key >> 6);
            uniq_j[pos] = static_cast<uint8_t>(key & 63);
        }
        base += tot;
        __syncthreads();
    }
    const int nu = base;
    if (tid == 0) n_cand[t] = nu;
    for (int u = tid; u < SEQ_MAXC; u += blockDim.x) cand_ids[t * SEQ_MAXC + u] = u < nu ? uniq_c[u] : -1;

    // gather / reduce: per candidate the prefix sums of q[j] . recon[c + j]
    const float* qt = qrows + t * L * D128;
    for (int u = warp; u < nu; u += 8) {
        const int64_t c = uniq_c[u];
        const bool owned = c >= owned_lo && c < owned_hi;
        const int64_t row = c - label_offset;
        int avail = 0;     // rows of recon that exist at c, c+1, ... (global extent, :224 slice clamps)
        if (owned) {
            const int64_t left = n_rows_global - c;
            avail = static_cast<int>(left < L ? left : L);
            const int64_t loc = n_rows_local - row;
            if (loc < avail) avail = static_cast<int>(loc < 0 ? 0 : loc);
        }
        const int terms = avail < lq ? avail : lq;
        float sum = 0.f;
        if (lane == 0) prefix[warp][0] = 0.f;
        for (int j = 0; j < terms; ++j) {
            const float4 qv = reinterpret_cast<const float4*>(qt + j * D128)[lane];
            sum += warp_dot128(qv, x32 + (row + j) * D128, lane);
            if (lane == 0) prefix[warp][j + 1] = sum;
        }
        __syncwarp();
        for (int li = lane; li < n_len; li += 32) {
            const int sl = seq_lens[li];
            const int lq_sl = sl < lq ? sl : lq;
            int m = sl < terms ? sl : terms;
            float sc = -INFINITY;
            if (owned && m > 0 && uniq_j[u] < lq_sl) sc = prefix[warp][m] / static_cast<float>(m);
            cand_scores[(t * n_len + li) * SEQ_MAXC + u] = sc;
        }
        __syncwarp();
    }
    for (int li = 0; li < n_len; ++li)
        for (int u = nu + tid; u < SEQ_MAXC; u += blockDim.x) cand_scores[(t * n_len + li) * SEQ_MAXC + u] = -INFINITY;
}

__global__ void __launch_bounds__(256)
seq_top_kernel(int n_len, const int64_t* __restrict__ cand_ids, const float* __restrict__ cand_scores,
               const int32_t* __restrict__ n_cand, int64_t* __restrict__ pred_ids, float* __restrict__ pred_scores) {
    __shared__ uint64_t red[8];
    __shared__ uint64_t win;
    const int64_t t = blockIdx.x / n_len;
    const int li = blockIdx.x % n_len;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nu = n_cand[t];
    const float* sc = cand_scores + (t * n_len + li) * SEQ_MAXC;
    uint64_t mine[SEQ_MAXC / 256];
#pragma unroll
    for (int r = 0; r < SEQ_MAXC / 256; ++r) {
        const int u = r * 256 + tid;
        uint64_t key = 0;
        if (u < nu) {
            const float s = sc[u];
            if (s > -INFINITY)
                key = (static_cast<uint64_t>(static_cast<uint32_t>(f2ord(s)) ^ 0x80000000u) << 32) |
                      static_cast<uint64_t>(0xFFFFFFFFu - static_cast<uint32_t>(u));
        }
        mine[r] = key;
    }
    for (int p = 0; p < SEQ_NPRED; ++p) {
        uint64_t best = 0;
#pragma unroll
        for (int r = 0; r < SEQ_MAXC / 256; ++r) best = mine[r] > best ? mine[r] : best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint64_t other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) red[warp] = best;
        __syncthreads();
        if (tid == 0) {
            uint64_t b = 0;
            for (int w = 0; w < 8; ++w) b = red[w] > b ? red[w] : b;
            win = b;
            const int64_t o = (t * n_len + li) * SEQ_NPRED + p;
            if (b != 0) {
                const uint32_t u = 0xFFFFFFFFu - static_cast<uint32_t>(b);
                pred_ids[o] = cand_ids[t * SEQ_MAXC + u];
                if (pred_scores) pred_scores[o] = ord2f(static_cast<int>(static_cast<uint32_t>(b >> 32) ^ 0x80000000u));
            } else {
                pred_ids[o] = -1;
                if (pred_scores) pred_scores[o] = -INFINITY;
            }
        }
        __syncthreads();
        const uint64_t w = win;
#pragma unroll
        for (int r = 0; r < SEQ_MAXC / 256; ++r)
            if (mine[r] == w) mine[r] = 0;
        __syncthreads();
    }
}


// merge W per-shard top-k lists (already holding global labels) into one: order by (distance, label)
__global__ void __launch_bounds__(128)
topk_merge_kernel(const float* __restrict__ D_all, const int64_t* __restrict__ I_all, int W, int64_t nq, int k,
                  float* __restrict__ D_out, int64_t* __restrict__ I_out) {
    __shared__ float d_s[2048];
    __shared__ int64_t i_s[2048];
    const int64_t q = blockIdx.x;
    const int n = W * k;
    int npow = 1;
    while (npow < n) npow <<= 1;
    for (int e = threadIdx.x; e < npow; e += blockDim.x) {
        float d = INFINITY;
        int64_t id = -1;
        if (e < n) {
            const int w = e / k, j = e % k;
            d = D_all[(static_cast<int64_t>(w) * nq + q) * k + j];
            id = I_all[(static_cast<int64_t>(w) * nq + q) * k + j];
        }
        if (id < 0) d = INFINITY;
        d_s[e] = d;
        i_s[e] = id < 0 ? LLONG_MAX : id;
    }
    __syncthreads();
    for (int kk = 2; kk <= npow; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npow; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const float da = d_s[i], db = d_s[ixj];
                    const int64_t ia = i_s[i], ib = i_s[ixj];
                    const bool a_after_b = (da > db) || (da == db && ia > ib);
                    const bool asc = (i & kk) == 0;
                    if (asc ? a_after_b : !a_after_b) {
                        d_s[i] = db; d_s[ixj] = da;
                        i_s[i] = ib; i_s[ixj] = ia;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const int64_t id = i_s[j];
        D_out[q * k + j] = d_s[j];
        I_out[q * k + j] = id == LLONG_MAX ? -1 : id;
    }
}

}  // namespace nafp

using namespace nafp;

extern "C" {

int nafp_seq_gather_dev(nafp_ctx* ctx, const float* q_dev, int64_t n_query_rows, const int64_t* test_ids_dev,
                        int64_t n_test, int32_t max_len, float* qrows_dev) {
    NAFP_REQUIRE(ctx && q_dev && test_ids_dev && qrows_dev && n_test >= 0 && max_len >= 1 && max_len <= SEQ_MAXL,
                 NAFP_ERR_INVALID, "nafp_seq_gather_dev: bad arguments (max_len <= %d)", SEQ_MAXL);
    if (n_test == 0) return NAFP_OK;
    const int64_t warps = n_test * max_len;
    seq_gather_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        q_dev, n_query_rows, test_ids_dev, n_test, max_len, qrows_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

int nafp_seq_cand_dev(nafp_index* idx, const float* qrows_dev, int64_t n_query_rows, const int64_t* test_ids_dev,
                      int64_t n_test, const int32_t* seq_lens_dev, int32_t n_len, int32_t max_len, int32_t k_probe,
                      const int64_t* I_dev, int64_t n_rows_global, int64_t owned_lo, int64_t owned_hi,
                      int64_t* cand_ids_dev, float* cand_scores_dev, int32_t* n_cand_dev) {
    NAFP_REQUIRE(idx && qrows_dev && test_ids_dev && seq_lens_dev && I_dev && cand_ids_dev && cand_scores_dev &&
                     n_cand_dev, NAFP_ERR_INVALID, "nafp_seq_cand_dev: NULL argument");
    NAFP_REQUIRE(max_len >= 1 && max_len <= SEQ_MAXL && k_probe >= 1 && max_len * k_probe <= SEQ_MAXC &&
                     n_len >= 1 && n_len <= 32, NAFP_ERR_INVALID,
                 "nafp_seq_cand_dev: need max_len <= %d, n_len <= 32, max_len*k_probe <= %d", SEQ_MAXL, SEQ_MAXC);
    NAFP_REQUIRE(idx->x32 != nullptr || idx->n == 0, NAFP_ERR_STATE, "nafp_seq_cand_dev: index holds no rows");
    if (n_test == 0) return NAFP_OK;
    nafp_ctx* ctx = idx->ctx;
    seq_cand_kernel<<<static_cast<unsigned>(n_test), 256, 0, ctx->stream>>>(
        qrows_dev, n_query_rows, test_ids_dev, seq_lens_dev, n_len, max_len, k_probe, I_dev, idx->x32, idx->n,
        n_rows_global, idx->label_offset, owned_lo, owned_hi, cand_ids_dev, cand_scores_dev, n_cand_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

int nafp_seq_top_dev(nafp_ctx* ctx, int64_t n_test, int32_t n_len, const int64_t* cand_ids_dev,
                     const float* cand_scores_dev, const int32_t* n_cand_dev, int64_t* pred_ids_dev,
                     float* pred_scores_dev) {
    NAFP_REQUIRE(ctx && cand_ids_dev && cand_scores_dev && n_cand_dev && pred_ids_dev && n_len >= 1,
                 NAFP_ERR_INVALID, "nafp_seq_top_dev: bad arguments");
    if (n_test == 0) return NAFP_OK;
    seq_top_kernel<<<static_cast<unsigned>(n_test * n_len), 256, 0, ctx->stream>>>(
        n_len, cand_ids_dev, cand_scores_dev, n_cand_dev, pred_ids_dev, pred_scores_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

int nafp_topk_merge_dev(nafp_ctx* ctx, const float* D_all_dev, const int64_t* I_all_dev, int32_t n_shards,
                        int64_t nq, int32_t k, float* D_out_dev, int64_t* I_out_dev) {
    NAFP_REQUIRE(ctx && D_all_dev && I_all_dev && D_out_dev && I_out_dev && n_shards >= 1 && k >= 1 &&
                     n_shards * k <= 2048, NAFP_ERR_INVALID, "nafp_topk_merge_dev: need n_shards*k <= 2048");
    if (nq == 0) return NAFP_OK;
    topk_merge_kernel<<<static_cast<unsigned>(nq), 128, 0, ctx->stream>>>(D_all_dev, I_all_dev, n_shards, nq, k,
                                                                          D_out_dev, I_out_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

int nafp_seq_match(nafp_index* idx, const float* q_host, int64_t n_query_rows, const int64_t* test_ids,
                   int64_t n_test, const int32_t* seq_lens, int32_t n_len, int32_t k_probe, int64_t* pred_ids_host,
                   float* pred_scores_host) {
    NAFP_REQUIRE(idx && q_host && test_ids && seq_lens && pred_ids_host && n_test >= 0 && n_len >= 1 &&
                     n_query_rows >= 0, NAFP_ERR_INVALID, "nafp_seq_match: bad arguments");
    if (n_test == 0) return NAFP_OK;
    int L = 0;
    for (int i = 0; i < n_len; ++i) {
        NAFP_REQUIRE(seq_lens[i] >= 1 && seq_lens[i] <= SEQ_MAXL, NAFP_ERR_INVALID,
                     "nafp_seq_match: sequence length %d outside [1,%d]", seq_lens[i], SEQ_MAXL);
        if (seq_lens[i] > L) L = seq_lens[i];
    }
    NAFP_REQUIRE(k_probe >= 1 && k_probe * L <= SEQ_MAXC && k_probe <= MAX_K, NAFP_ERR_INVALID,
                 "nafp_seq_match: k_probe*max_len must be <= %d", SEQ_MAXC);
    nafp_ctx* ctx = idx->ctx;
    NAFP_CUDA(cudaSetDevice(ctx->device));
    const int64_t rows = n_test * L;
    // one arena for all temporaries
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_q = take(static_cast<size_t>(n_query_rows) * D128 * 4);
    const size_t o_ids = take(static_cast<size_t>(n_test) * 8);
    const size_t o_sl = take(static_cast<size_t>(n_len) * 4);
    const size_t o_qrows = take(static_cast<size_t>(rows) * D128 * 4);
    const size_t o_D = take(static_cast<size_t>(rows) * k_probe * 4);
    const size_t o_I = take(static_cast<size_t>(rows) * k_probe * 8);
    const size_t o_cid = take(static_cast<size_t>(n_test) * SEQ_MAXC * 8);
    const size_t o_csc = take(static_cast<size_t>(n_test) * n_len * SEQ_MAXC * 4);
    const size_t o_nc = take(static_cast<size_t>(n_test) * 4);
    const size_t o_pid = take(static_cast<size_t>(n_test) * n_len * SEQ_NPRED * 8);
    const size_t o_psc = take(static_cast<size_t>(n_test) * n_len * SEQ_NPRED * 4);
    NAFP_TRY(ensure_dev(ctx, &ctx->stage_dev, &ctx->stage_dev_bytes, static_cast<int64_t>(off)));
    uint8_t* base = static_cast<uint8_t*>(ctx->stage_dev);
    float* q_dev = reinterpret_cast<float*>(base + o_q);
    int64_t* ids_dev = reinterpret_cast<int64_t*>(base + o_ids);
    int32_t* sl_dev = reinterpret_cast<int32_t*>(base + o_sl);
    float* qrows = reinterpret_cast<float*>(base + o_qrows);
    float* Dd = reinterpret_cast<float*>(base + o_D);
    int64_t* Id = reinterpret_cast<int64_t*>(base + o_I);
    int64_t* cid = reinterpret_cast<int64_t*>(base + o_cid);
    float* csc = reinterpret_cast<float*>(base + o_csc);
    int32_t* nc = reinterpret_cast<int32_t*>(base + o_nc);
    int64_t* pid = reinterpret_cast<int64_t*>(base + o_pid);
    float* psc = reinterpret_cast<float*>(base + o_psc);
    NAFP_CUDA(cudaMemcpyAsync(q_dev, q_host, static_cast<size_t>(n_query_rows) * D128 * 4, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_CUDA(cudaMemcpyAsync(ids_dev, test_ids, static_cast<size_t>(n_test) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_CUDA(cudaMemcpyAsync(sl_dev, seq_lens, static_cast<size_t>(n_len) * 4, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_TRY(nafp_seq_gather_dev(ctx, q_dev, n_query_rows, ids_dev, n_test, L, qrows));
    NAFP_TRY(nafp_index_search_dev(idx, qrows, rows, k_probe, Dd, Id));
    const int64_t n_search = idx->search_rows >= 0 && idx->search_rows < idx->n ? idx->search_rows : idx->n;
    NAFP_TRY(nafp_seq_cand_dev(idx, qrows, n_query_rows, ids_dev, n_test, sl_dev, n_len, L, k_probe, Id,
                               idx->label_offset + idx->n, idx->label_offset, idx->label_offset + n_search, cid, csc, nc));
    NAFP_TRY(nafp_seq_top_dev(ctx, n_test, n_len, cid, csc, nc, pid, psc));
    NAFP_CUDA(cudaMemcpyAsync(pred_ids_host, pid, static_cast<size_t>(n_test) * n_len * SEQ_NPRED * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (pred_scores_host)
        NAFP_CUDA(cudaMemcpyAsync(pred_scores_host, psc, static_cast<size_t>(n_test) * n_len * SEQ_NPRED * 4, cudaMemcpyDeviceToHost, ctx->stream));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    return NAFP_OK;
}

}  // extern "C"
