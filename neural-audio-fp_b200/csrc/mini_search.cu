// In-training mini search: the reference's model/utils/mini_search_subroutines.py on B200.
//
//   pairdist_kernel    pairwise_distances_for_eval (:29-90): all query x db squared-L2 distances (or dot products)
//                      as one fp32 tile GEMM on CUDA cores -- the matrices are a few thousand rows, the work is
//                      exact fp32 (ranks are compared against an fp64 oracle) and far too small for a tensor-core
//                      pipeline to pay off.
//   conv_eye_kernel    conv_eye_func (:93-120): "convolution with an identity kernel" = sums over s consecutive
//                      diagonal elements, conv[a,i,j] = sum_t dist[a,i+t,j+t] ('valid' padding).
//   mini_rank_kernel   mini_search_eval (:123-236) without materialising conv or sorting it: the rank of the
//                      ground-truth item gt in argsort(conv[a,i,:]) is the number of items that sort before it,
//                      so one block per (scope, augmentation, target) computes conv[a,i,gt] and counts.  Top-1/3/10
//                      accuracy and the mean rank follow from the ranks on the host.
// All of it is HBM/L2-resident gather-reduce work: nQ x nD x s fp32 loads per scope.
#include <cmath>
#include <vector>

#include "common.h"

namespace nafp {

constexpr int PD_TILE = 64;      // 64 x 64 outputs per block, 4 x 4 per thread
constexpr int PD_K = 16;

// squared norms of the rows of x (n, d): one warp per row
__global__ void __launch_bounds__(256)
rownorm_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    float s = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float v = x[r * d + k];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[r] = s;
}

// q: (nQ, nAug, d) row (i, a) at (i * nAug + a) * d; db: (nD, d); out: (nAug, nQ, nD)
// mode 0: dot product; 1: max(|q|^2 + |x|^2 - 2 q.x, 0); 2: its square root with the reference's zero mask
__global__ void __launch_bounds__(256)
pairdist_kernel(const float* __restrict__ q, const float* __restrict__ db, const float* __restrict__ qsq,
                const float* __restrict__ dsq, int nQ, int nAug, int nD, int d, int mode, float* __restrict__ out) {
    __shared__ float As[PD_K][PD_TILE + 1], Bs[PD_K][PD_TILE + 1];
    const int a = blockIdx.z;
    const int i0 = blockIdx.y * PD_TILE, j0 = blockIdx.x * PD_TILE;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < d; k0 += PD_K) {
        for (int e = threadIdx.x; e < PD_TILE * PD_K; e += 256) {
            const int r = e / PD_K, k = e % PD_K;
            const int i = i0 + r, j = j0 + r;
            As[k][r] = (i < nQ && k0 + k < d) ? q[(static_cast<int64_t>(i) * nAug + a) * d + k0 + k] : 0.f;
            Bs[k][r] = (j < nD && k0 + k < d) ? db[static_cast<int64_t>(j) * d + k0 + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PD_K; ++k) {
            float av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                av[u] = As[k][ty * 4 + u];
                bv[u] = Bs[k][tx * 4 + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + ty * 4 + u;
        if (i >= nQ) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int j = j0 + tx * 4 + v;
            if (j >= nD) continue;
            float r = acc[u][v];
            if (mode != 0) {
                // ||a||^2 + ||b||^2 - 2 a.b, clamped at 0 (mini_search_subroutines.py:74-79)
                r = fmaxf(qsq[static_cast<int64_t>(i) * nAug + a] + dsq[j] - 2.0f * r, 0.f);
                if (mode == 2) r = r == 0.f ? 0.f : sqrtf(r);      // (:82-86): sqrt(d + 1e-16 mask) * (1 - mask)
            }
            out[(static_cast<int64_t>(a) * nQ + i) * nD + j] = r;
        }
    }
}

// x: (nA, nQ, nD) -> out: (nA, nQ - s + 1, nD - s + 1), out[a,i,j] = sum_t x[a, i+t, j+t]
__global__ void __launch_bounds__(256)
conv_eye_kernel(const float* __restrict__ x, int nQ, int nD, int s, float* __restrict__ out) {
    const int oQ = nQ - s + 1, oD = nD - s + 1;
    const int a = blockIdx.z, i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= oD) return;
    const float* base = x + (static_cast<int64_t>(a) * nQ + i) * nD + j;
    float acc = 0.f;
    for (int t = 0; t < s; ++t) acc += base[static_cast<int64_t>(t) * (nD + 1)];
    out[(static_cast<int64_t>(a) * oQ + i) * oD + j] = acc;
}

// block (target i, augmentation a, scope si): rank of gt = i + gt_off among conv[a,i,0..oD), or -1 when gt is
// outside the 'valid' range (np.where finds nothing there, :195-197).  Items that sort before gt: strictly smaller
// (argmin) / strictly larger (argmax) values; exact ties are ordered by index the way a stable ascending sort --
// reversed for argmax (:189-191) -- orders them.
__global__ void __launch_bounds__(256)
mini_rank_kernel(const float* __restrict__ dist, int nQ, int nD, const int32_t* __restrict__ scopes, int argmax,
                 int64_t gt_off, int max_targets, int32_t* __restrict__ rank_out) {
    __shared__ int cnt_s[8];
    const int s = scopes[blockIdx.z];
    const int oQ = nQ - s + 1, oD = nD - s + 1;
    const int a = blockIdx.y, i = blockIdx.x;
    int32_t* dst = rank_out + (static_cast<int64_t>(blockIdx.z) * gridDim.y + a) * max_targets + i;
    if (i >= oQ) return;
    const int64_t gt = i + gt_off;
    if (gt < 0 || gt >= oD || oD <= 0) {
        if (threadIdx.x == 0) *dst = -1;
        return;
    }
    const float* row = dist + (static_cast<int64_t>(a) * nQ + i) * nD;
    float vgt = 0.f;
    for (int t = 0; t < s; ++t) vgt += row[static_cast<int64_t>(t) * (nD + 1) + gt];
    int cnt = 0;
    for (int j = threadIdx.x; j < oD; j += blockDim.x) {
        float v = 0.f;
        for (int t = 0; t < s; ++t) v += row[static_cast<int64_t>(t) * (nD + 1) + j];
        if (argmax) cnt += (v > vgt) || (v == vgt && j > gt);
        else        cnt += (v < vgt) || (v == vgt && j < gt);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) cnt_s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += cnt_s[w];
        *dst = tot;
    }
}

struct DevBuf {            // device temporaries released on every return path
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) {
        NAFP_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
        return NAFP_OK;
    }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

// uploads q / db and fills dist_dev (nAug, nQ, nD)
static int pairwise_dev(nafp_ctx* ctx, const float* q_host, const float* db_host, int64_t nQ, int64_t nAug, int64_t nD,
                        int64_t d, int mode, DevBuf& qd, DevBuf& dbd, DevBuf& norms, float* dist_dev) {
    cudaStream_t st = ctx->stream;
    NAFP_TRY(qd.alloc(static_cast<size_t>(nQ * nAug * d) * sizeof(float)));
    NAFP_TRY(dbd.alloc(static_cast<size_t>(nD * d) * sizeof(float)));
    NAFP_TRY(norms.alloc(static_cast<size_t>(nQ * nAug + nD) * sizeof(float)));
    NAFP_CUDA(cudaMemcpyAsync(qd.p, q_host, static_cast<size_t>(nQ * nAug * d) * sizeof(float), cudaMemcpyHostToDevice, st));
    NAFP_CUDA(cudaMemcpyAsync(dbd.p, db_host, static_cast<size_t>(nD * d) * sizeof(float), cudaMemcpyHostToDevice, st));
    float* qsq = norms.as<float>();
    float* dsq = qsq + nQ * nAug;
    rownorm_kernel<<<static_cast<unsigned>((nQ * nAug * 32 + 255) / 256), 256, 0, st>>>(qd.as<float>(), nQ * nAug, static_cast<int>(d), qsq);
    rownorm_kernel<<<static_cast<unsigned>((nD * 32 + 255) / 256), 256, 0, st>>>(dbd.as<float>(), nD, static_cast<int>(d), dsq);
    const dim3 grid(static_cast<unsigned>((nD + PD_TILE - 1) / PD_TILE), static_cast<unsigned>((nQ + PD_TILE - 1) / PD_TILE),
                    static_cast<unsigned>(nAug));
    pairdist_kernel<<<grid, 256, 0, st>>>(qd.as<float>(), dbd.as<float>(), qsq, dsq, static_cast<int>(nQ), static_cast<int>(nAug),
                                          static_cast<int>(nD), static_cast<int>(d), mode, dist_dev);
    ctx->launches += 3;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

static bool mini_shapes_ok(int64_t nQ, int64_t nAug, int64_t nD, int64_t d) {
    return nQ >= 1 && nAug >= 1 && nD >= 1 && d >= 1 && nQ <= 65535 * 64ll && nAug <= 65535 && nD < (1ll << 31) &&
           nQ * nAug * nD < (1ll << 40);
}

}  // namespace nafp

using namespace nafp;

extern "C" {

int nafp_pairwise_dists_host(nafp_ctx* ctx, const float* q_host, const float* db_host, int64_t n_q, int64_t n_aug,
                             int64_t n_db, int64_t d, int32_t return_dotprod, int32_t squared, float* out_host) {
    NAFP_RANGE("nafp_pairwise_dists_host");
    NAFP_REQUIRE(ctx && q_host && db_host && out_host && mini_shapes_ok(n_q, n_aug, n_db, d), NAFP_ERR_INVALID,
                 "nafp_pairwise_dists_host: bad arguments");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    DevBuf qd, dbd, norms, dist;
    const size_t bytes = static_cast<size_t>(n_aug * n_q * n_db) * sizeof(float);
    NAFP_TRY(dist.alloc(bytes));
    NAFP_TRY(pairwise_dev(ctx, q_host, db_host, n_q, n_aug, n_db, d, return_dotprod ? 0 : (squared ? 1 : 2), qd, dbd, norms,
                          dist.as<float>()));
    NAFP_CUDA(cudaMemcpyAsync(out_host, dist.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    return NAFP_OK;
}

int nafp_conv_eye_host(nafp_ctx* ctx, const float* x_host, int64_t n_aug, int64_t n_q, int64_t n_db, int32_t s,
                       float* out_host) {
    NAFP_RANGE("nafp_conv_eye_host");
    NAFP_REQUIRE(ctx && x_host && out_host && s >= 1 && mini_shapes_ok(n_q, n_aug, n_db, 1) && n_q <= 65535, NAFP_ERR_INVALID,
                 "nafp_conv_eye_host: bad arguments");
    NAFP_REQUIRE(s <= n_q && s <= n_db, NAFP_ERR_INVALID, "nafp_conv_eye_host: scope %d larger than the %lld x %lld matrix",
                 (int)s, (long long)n_q, (long long)n_db);
    NAFP_CUDA(cudaSetDevice(ctx->device));
    DevBuf x, o;
    const int64_t oQ = n_q - s + 1, oD = n_db - s + 1;
    NAFP_TRY(x.alloc(static_cast<size_t>(n_aug * n_q * n_db) * sizeof(float)));
    NAFP_TRY(o.alloc(static_cast<size_t>(n_aug * oQ * oD) * sizeof(float)));
    NAFP_CUDA(cudaMemcpyAsync(x.p, x_host, static_cast<size_t>(n_aug * n_q * n_db) * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid(static_cast<unsigned>((oD + 255) / 256), static_cast<unsigned>(oQ), static_cast<unsigned>(n_aug));
    conv_eye_kernel<<<grid, 256, 0, ctx->stream>>>(x.as<float>(), static_cast<int>(n_q), static_cast<int>(n_db), s, o.as<float>());
    ctx->launches += 1;
    NAFP_CUDA(cudaGetLastError());
    NAFP_CUDA(cudaMemcpyAsync(out_host, o.p, static_cast<size_t>(n_aug * oQ * oD) * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    return NAFP_OK;
}

int nafp_mini_search_host(nafp_ctx* ctx, const float* q_host, const float* db_host, int64_t n_q, int64_t n_aug, int64_t n_db,
                          int64_t d, const int32_t* scopes, int32_t n_scopes, int32_t argmax, int64_t gt_id_offset,
                          double* top1, double* top3, double* top10, double* mean_rank) {
    NAFP_RANGE("nafp_mini_search_host");
    NAFP_REQUIRE(ctx && q_host && db_host && scopes && n_scopes >= 1 && top1 && top3 && top10 && mean_rank &&
                     mini_shapes_ok(n_q, n_aug, n_db, d) && n_q <= (1 << 30),
                 NAFP_ERR_INVALID, "nafp_mini_search_host: bad arguments");
    for (int i = 0; i < n_scopes; ++i)
        NAFP_REQUIRE(scopes[i] >= 1 && scopes[i] <= n_q && scopes[i] <= n_db, NAFP_ERR_INVALID,
                     "nafp_mini_search_host: scope %d does not fit the %lld x %lld distance matrix", (int)scopes[i],
                     (long long)n_q, (long long)n_db);
    NAFP_CUDA(cudaSetDevice(ctx->device));
    DevBuf qd, dbd, norms, dist, sc, rk;
    NAFP_TRY(dist.alloc(static_cast<size_t>(n_aug * n_q * n_db) * sizeof(float)));
    NAFP_TRY(pairwise_dev(ctx, q_host, db_host, n_q, n_aug, n_db, d, argmax ? 0 : 1, qd, dbd, norms, dist.as<float>()));
    NAFP_TRY(sc.alloc(n_scopes * sizeof(int32_t)));
    const size_t n_rank = static_cast<size_t>(n_scopes) * n_aug * n_q;
    NAFP_TRY(rk.alloc(n_rank * sizeof(int32_t)));
    NAFP_CUDA(cudaMemcpyAsync(sc.p, scopes, n_scopes * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    NAFP_CUDA(cudaMemsetAsync(rk.p, 0xff, n_rank * sizeof(int32_t), ctx->stream));
    const dim3 grid(static_cast<unsigned>(n_q), static_cast<unsigned>(n_aug), static_cast<unsigned>(n_scopes));
    mini_rank_kernel<<<grid, 256, 0, ctx->stream>>>(dist.as<float>(), static_cast<int>(n_q), static_cast<int>(n_db), sc.as<int32_t>(),
                                                    argmax ? 1 : 0, gt_id_offset, static_cast<int>(n_q), rk.as<int32_t>());
    ctx->launches += 1;
    NAFP_CUDA(cudaGetLastError());
    std::vector<int32_t> ranks(n_rank);
    NAFP_CUDA(cudaMemcpyAsync(ranks.data(), rk.p, n_rank * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    // accuracies and mean rank exactly as the reference accumulates them (:193-213): sums over augmentations / n_augs,
    // then / n_targets; a ground truth outside the 'valid' range contributes nothing
    for (int si = 0; si < n_scopes; ++si) {
        const int64_t n_targets = n_q - scopes[si] + 1;
        double r_sum = 0.0, c1 = 0.0, c3 = 0.0, c10 = 0.0;
        for (int64_t a = 0; a < n_aug; ++a)
            for (int64_t i = 0; i < n_targets; ++i) {
                const int32_t r = ranks[(static_cast<size_t>(si) * n_aug + a) * n_q + i];
                if (r < 0) continue;
                r_sum += r;
                c1 += r == 0;
                c3 += r < 3;
                c10 += r < 10;
            }
        const double den = static_cast<double>(n_aug) * static_cast<double>(n_targets);
        mean_rank[si] = r_sum / den;
        top1[si] = 100.0 * c1 / den;
        top3[si] = 100.0 * c3 / den;
        top10[si] = 100.0 * c10 / den;
    }
    return NAFP_OK;
}

}  // extern "C"
