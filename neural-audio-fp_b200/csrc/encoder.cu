// FingerPrinter encoder: the reference's model/fp/nnfp.py on B200.
//
//   conv0_a            1x3 conv with C_in = 1 (K = 3): CUDA cores, fused with the log-mel max
//                      subtraction / clamp, bias, ELU and LayerNorm statistics.
//   conv_gemm_kernel   the other 15 separable convolutions as implicit GEMMs on tcgen05:
//                      M = 128 output positions (NHWC rows), N = C_out tile, K = 3 taps x C_in.
//                      The A operand of each (tap, 64-channel block) is ONE TMA box of the previous
//                      layer's normalised fp16 activation: stride-2 axes are split into
//                      (parity, half) dimensions of the tensor map, TF 'SAME' zero padding is TMA
//                      out-of-bounds fill.  Persistent CTAs, 4-stage smem ring, fp32 accumulators
//                      double-buffered in TMEM, 16 epilogue warps: bias + ELU + per-sample
//                      sum / sum-of-squares slots (LayerNorm over (F,T,C)) + fp16 store; the 128-channel
//                      layers keep their weights resident in shared memory.
//   ln_apply_kernel    reduces the slots, (y - mean) * rstd * gamma[f,t,c] + beta[f,t,c]  ->  fp16 operand of the next conv
//   divenc_kernel      last LayerNorm + divide-and-encode head (128 x [8->32 ELU, 32->1]) + L2 norm.
// fp16 operands / fp32 accumulation: measured against the fp64 oracle the fingerprints agree to
// ~1e-4 (gate: cosine >= 0.9999, max abs <= 1e-3).
#include <cuda_fp16.h>

#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace nafp {

int logmel_run(nafp_ctx* ctx, const void* x_dev, bool pcm16, int64_t n_seg, int64_t group_size, float* mel_dev,
               bool finish, const int32_t** gmax_out, const int64_t* seg_off, const int32_t* seg_valid);

constexpr int ENC_LAYERS = 16;
constexpr int ENC_CHUNK_MAX = 4000;    // most segments of one encoder pass (2.3 MB of activations each): at 1,000 the
                                       // six last layers have fewer tiles than the chip has SMs
constexpr int ENC_CHUNK_MIN = 1000;    // the activation arena starts here and grows on demand (encoder_reserve)
constexpr int EMB = 128;
constexpr float LN_EPS = 1e-3f;        // Keras LayerNormalization default
constexpr float L2_EPS = 1e-12f;       // tf.math.l2_normalize default

struct ConvGeom {
    int axis, stride, pad_lo;          // axis 0 = time (1x3), 1 = frequency (3x1)
    int f_in, t_in, c_in, f_out, t_out, c_out;
    int ms;                            // output positions per segment
    int mode;                          // TMA addressing mode, see conv_gemm_kernel
    int tap_lo, tap_hi;                // taps that touch real data (others are all padding)
    int bt, bf, bb;                    // box extents in output positions: time, freq, segments
    int nt;                            // N tile
    int ksplit;                        // 3 = this layer's input is stored [hi | hi | lo] per position and its
                                       // weights [hi | lo | hi] per tap (hi + lo = the fp32 value to 2^-22):
                                       // the same implicit GEMM then computes hi.hi + hi.lo + lo.hi
    int osplit;                        // split factor of this layer's stored output (= ksplit of its consumer)
};
constexpr int ENC_SPLIT_FROM = 6;      // layers >= L4a: 25 % of the flops; removes ~35 % of the fingerprint's worst-case fp16 error
constexpr int ENC_Y32_FROM = 6;        // the same layers store their pre-LayerNorm output in fp32 (<= 64 KB per segment)

struct ConvParams {
    int m_total, ms, c_in, c_out, nt, n_ntiles, n_mtiles;
    int mode, pad_lo, tap_lo, tap_hi, kb_per_tap, tps, bf, bb;
    int groups;      // 32-row groups per segment (>= 1): slots of the LayerNorm partial sums
    int y32;         // 1 = the pre-LayerNorm output is stored in fp32 (split-precision layers), 0 = fp16
};

struct EncoderState {
    ConvGeom g[ENC_LAYERS];
    float* w0 = nullptr;                       // conv0_a kernel [3][128]
    __half* wt[ENC_LAYERS] = {};               // [c_out][3*c_in] fp16, K-major
    float* bias[ENC_LAYERS] = {};
    float* ln_g[ENC_LAYERS] = {};
    float* ln_b[ENC_LAYERS] = {};
    float *dw1 = nullptr, *db1 = nullptr, *dw2 = nullptr, *db2 = nullptr;
    __half* y = nullptr;                       // pre-LayerNorm scratch, largest layer
    __half* x[ENC_LAYERS] = {};                // normalised activations
    int cap = 0;                               // segments the arena below holds (ENC_CHUNK_MIN .. ENC_CHUNK_MAX)
    float* stats = nullptr;                    // [ENC_LAYERS][cap][2]
    float* part = nullptr;                     // [cap][64 x parts][2] partial LayerNorm sums of the running layer
    float* mel = nullptr;                      // (cap, 256, 32) for the fused entry points
    void* xin[2] = {nullptr, nullptr};         // (cap, 8000) staging for the *_host entry points, double-buffered:
    cudaStream_t copy_stream = nullptr;        // the upload of pass i+1 runs on copy_stream under the kernels of pass i
    cudaEvent_t ev_up[2] = {nullptr, nullptr}; // xin[b] has landed
    cudaEvent_t ev_free[2] = {nullptr, nullptr};   // the last pass that read xin[b] has finished
    float* emb = nullptr;                      // (cap, 128)
    float* raw = nullptr;                      // (cap, 128) head outputs before the L2 normalisation
    CUtensorMap tmA[ENC_LAYERS], tmB[ENC_LAYERS];
    bool weights = false;
    int64_t last_n = 0;                        // segments of the last pass (activation probe)
};

static const int kHidden[8] = {128, 128, 256, 256, 512, 512, 1024, 1024};      // nnfp.py:193
static const int kStrideT[8] = {2, 2, 2, 2, 1, 2, 1, 2};                        // nnfp.py:194-197 (1x3 conv)

static void same_pad(int n_in, int k, int s, int* n_out, int* lo) {
    *n_out = (n_in + s - 1) / s;
    int total = (*n_out - 1) * s + k - n_in;
    if (total < 0) total = 0;
    *lo = total / 2;
}

static void build_geometry(ConvGeom* g) {
    int f = 256, t = 32, c = 1;
    for (int i = 0; i < 8; ++i) {
        for (int half = 0; half < 2; ++half) {
            ConvGeom& L = g[2 * i + half];
            L.axis = half;
            L.stride = half == 0 ? kStrideT[i] : 2;
            L.f_in = f; L.t_in = t; L.c_in = c; L.c_out = kHidden[i];
            if (half == 0) { same_pad(t, 3, L.stride, &L.t_out, &L.pad_lo); L.f_out = f; }
            else           { same_pad(f, 3, L.stride, &L.f_out, &L.pad_lo); L.t_out = t; }
            L.ms = L.f_out * L.t_out;
            const int n_in = half == 0 ? L.t_in : L.f_in, n_out = half == 0 ? L.t_out : L.f_out;
            L.tap_lo = 3; L.tap_hi = -1;
            for (int tap = 0; tap < 3; ++tap) {
                bool any = false;
                for (int o = 0; o < n_out; ++o) {
                    const int src = o * L.stride + tap - L.pad_lo;
                    if (src >= 0 && src < n_in) any = true;
                }
                if (any) { if (tap < L.tap_lo) L.tap_lo = tap; if (tap > L.tap_hi) L.tap_hi = tap; }
            }
            if (n_out == 1 || L.stride == 1) L.mode = half == 0 ? 0 : 3;
            else L.mode = half == 0 ? 1 : 2;
            L.bt = L.t_out;
            L.bb = L.ms >= 128 ? 1 : 128 / L.ms;
            L.bf = L.ms >= 128 ? 128 / L.t_out : L.f_out;
            L.nt = L.c_out < 256 ? L.c_out : 256;
            f = L.f_out; t = L.t_out; c = L.c_out;
        }
    }
    for (int l = 0; l < 16; ++l) g[l].ksplit = l >= ENC_SPLIT_FROM ? 3 : 1;
    for (int l = 0; l < 16; ++l) g[l].osplit = l + 1 < 16 ? g[l + 1].ksplit : 3;     // the head reads hi + lo too
}

// ELU(v) = v > 0 ? v : e^v - 1.  Only the v <= 0 side reaches the exponential, where e^v is in (0, 1]: the bare
// MUFU ex2 (flush-to-zero) is exact enough (2 ulp, absolute error 2^-22 after the -1) and saves the
// denormal-range fix-up of __expf -- 5 of the ~15 instructions per output element of the conv epilogues.
__device__ __forceinline__ float elu(float v) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
    return v > 0.f ? v : e - 1.f;
}

// ------------------------------------------------------------------------------------------
// conv0_a: (B,256,32) fp32 log-mel -> (B,256,16,128) normalised fp16, stride 2 in time, pad (0,1).
// K = 3: CUDA cores; fused with the log-mel max subtraction / clamp, bias, ELU and the LayerNorm.
// ------------------------------------------------------------------------------------------
struct Conv0Lane {
    float w[3][4], bia[4];
};
__device__ __forceinline__ void conv0_load_lane(const float* __restrict__ w0, const float* __restrict__ b0, int lane,
                                                Conv0Lane& L) {
#pragma unroll
    for (int tap = 0; tap < 3; ++tap) {
        const float4 v = reinterpret_cast<const float4*>(w0 + tap * 128)[lane];
        L.w[tap][0] = v.x; L.w[tap][1] = v.y; L.w[tap][2] = v.z; L.w[tap][3] = v.w;
    }
    const float4 v = reinterpret_cast<const float4*>(b0)[lane];
    L.bia[0] = v.x; L.bia[1] = v.y; L.bia[2] = v.z; L.bia[3] = v.w;
}
// L1a never materialises its pre-LayerNorm output (1 MB fp16 per segment would make an HBM round trip):
//   conv0_stats_kernel  ELU(conv) of every output, reduced to (sum, sum of squares) slots -- no store;
//   conv0_ln_kernel     the same values recomputed, normalised with gamma/beta and stored once.
// The exponentials are cheaper than the 2 GB of traffic they replace.
// Both: a warp per frequency row; the row's 32 log-mel values are one coalesced load, the three taps of
// every output position come from warp shuffles, the 16 positions of the row are unrolled; lane owns 4 channels.
__device__ __forceinline__ void conv0_row(float v, const Conv0Lane& L, int tp, float (&o)[4]) {
    const float x0 = __shfl_sync(0xffffffffu, v, 2 * tp);
    const float x1 = __shfl_sync(0xffffffffu, v, 2 * tp + 1);
    const float x2 = tp < 15 ? __shfl_sync(0xffffffffu, v, (2 * tp + 2) & 31) : 0.f;      // SAME padding (0, 1)
#pragma unroll
    for (int c = 0; c < 4; ++c) o[c] = elu(L.bia[c] + x0 * L.w[0][c] + x1 * L.w[1][c] + x2 * L.w[2][c]);
}

// grid (8, n_seg): block (bx, seg) covers frequency rows 32 bx .. 32 bx + 31 of one segment
__global__ void __launch_bounds__(256)
conv0_stats_kernel(const float* __restrict__ mel, const int32_t* __restrict__ gmax, int64_t group_size, int n_seg,
                   const float* __restrict__ w0, const float* __restrict__ b0, float* __restrict__ part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = blockIdx.y;
    Conv0Lane L;
    conv0_load_lane(w0, b0, lane, L);
    const bool raw = gmax != nullptr;
    const float sub = raw ? ord2f(gmax[seg / group_size]) : 0.f;
    const float* m = mel + static_cast<int64_t>(seg) * 8192;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        const int f = blockIdx.x * 32 + r * 8 + warp;
        float v = __ldg(m + f * 32 + lane);
        if (raw) v = fmaxf(v - sub, -80.f);                 // "- batch max, clamp -80" (melspectrogram.py:108-109)
#pragma unroll
        for (int tp = 0; tp < 16; ++tp) {
            float o[4];
            conv0_row(v, L, tp, o);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                s1 += o[c];
                s2 += o[c] * o[c];
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0)      // one slot per (segment, block, warp): summed in fixed order by conv0_ln_kernel
        reinterpret_cast<float2*>(part)[static_cast<int64_t>(seg) * 64 + blockIdx.x * 8 + warp] = make_float2(s1, s2);
}

// grid (8, ceil(n_seg / 8)): block (bx, by) covers rows 32 bx .. 32 bx + 31 of segments 8 by .. 8 by + 7, so that
// gamma / beta (8 B per element, shared by all segments) come from L1 after the first segment
constexpr int CONV0_SEGS = 8;
__global__ void __launch_bounds__(256)
conv0_ln_kernel(const float* __restrict__ mel, const int32_t* __restrict__ gmax, int64_t group_size, int n_seg,
                const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ part,
                float* __restrict__ stats_out, const float* __restrict__ gamma, const float* __restrict__ beta,
                __half* __restrict__ x) {
    __shared__ float2 mr_s[CONV0_SEGS];          // (mean, rstd) of the block's segments
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg0 = blockIdx.y * CONV0_SEGS;
    constexpr int PER_SEG = 256 * 16 * 128;
    {
        const int seg = seg0 + warp;             // 8 warps = CONV0_SEGS segments
        if (seg < n_seg) {
            const float2* p = reinterpret_cast<const float2*>(part) + static_cast<int64_t>(seg) * 64;
            double a = 0.0, b = 0.0;
            for (int i = lane; i < 64; i += 32) {
                const float2 v = p[i];
                a += v.x;
                b += v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) {
                const float inv_n = 1.f / static_cast<float>(PER_SEG);
                const float s1 = static_cast<float>(a), s2 = static_cast<float>(b);
                const float mean = s1 * inv_n;
                const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
                mr_s[warp] = make_float2(mean, rsqrtf(var + LN_EPS));
                if (blockIdx.x == 0 && stats_out) {
                    stats_out[2 * seg] = s1;
                    stats_out[2 * seg + 1] = s2;
                }
            }
        }
    }
    __syncthreads();
    Conv0Lane L;
    conv0_load_lane(w0, b0, lane, L);
    const bool raw = gmax != nullptr;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
        const int f = blockIdx.x * 32 + r * 8 + warp;
        const float4* g_row = reinterpret_cast<const float4*>(gamma + static_cast<int64_t>(f) * (16 * 128)) + lane;
        const float4* b_row = reinterpret_cast<const float4*>(beta + static_cast<int64_t>(f) * (16 * 128)) + lane;
#pragma unroll 1
        for (int k = 0; k < CONV0_SEGS; ++k) {
            const int seg = seg0 + k;
            if (seg >= n_seg) break;
            float v = __ldg(mel + static_cast<int64_t>(seg) * 8192 + f * 32 + lane);
            if (raw) v = fmaxf(v - ord2f(gmax[seg / group_size]), -80.f);
            const float mean = mr_s[k].x, rstd = mr_s[k].y;
            uint2* orow = reinterpret_cast<uint2*>(x + static_cast<int64_t>(seg) * PER_SEG + static_cast<int64_t>(f) * (16 * 128)) + lane;
#pragma unroll
            for (int tp = 0; tp < 16; ++tp) {
                float o[4];
                conv0_row(v, L, tp, o);
                const float4 g = __ldg(g_row + tp * 32), b = __ldg(b_row + tp * 32);
                const float y0 = (o[0] - mean) * rstd * g.x + b.x, y1 = (o[1] - mean) * rstd * g.y + b.y;
                const float y2 = (o[2] - mean) * rstd * g.z + b.z, y3 = (o[3] - mean) * rstd * g.w + b.w;
                __half2 h0 = __floats2half2_rn(y0, y1), h1 = __floats2half2_rn(y2, y3);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&h0);
                pk.y = *reinterpret_cast<uint32_t*>(&h1);
                orow[tp * 32] = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// implicit-GEMM separable convolution on tcgen05
// ------------------------------------------------------------------------------------------
constexpr int CONV_STAGES = 4;
constexpr int CONV_EPI_WARPS = 16;         // 4 per TMEM lane quadrant: each takes a quarter of the tile's columns
constexpr int CONV_EPI_PARTS = CONV_EPI_WARPS / 4;
constexpr int CONV_THREADS = (2 + CONV_EPI_WARPS) * 32;   // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2.. epilogue
constexpr int CONV_A_BYTES = 128 * 128;    // 128 rows x 64 fp16
constexpr int CONV_B_BYTES_MAX = 256 * 128;
constexpr int CONV_SMEM = CONV_STAGES * (CONV_A_BYTES + CONV_B_BYTES_MAX) + 1024 * 4 + 256 + 1024;

struct ConvBars {
    uint64_t full[CONV_STAGES];
    uint64_t empty[CONV_STAGES];
    uint64_t tfull[2];
    uint64_t tempty[2];
    uint64_t bfull;            // resident-weights mode: all K blocks of the CTA's N tile have landed
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const ConvParams p, const float* __restrict__ bias, __half* __restrict__ y,
                 float* __restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int b_bytes = p.nt * 128;
    uint8_t* a_s = smem;                                        // [stage][128][128 B]
    uint8_t* b_s = smem + CONV_STAGES * CONV_A_BYTES;           // [stage][nt][128 B]
    float* bias_s = reinterpret_cast<float*>(b_s + CONV_STAGES * CONV_B_BYTES_MAX);   // [c_out <= 1024]
    ConvBars* bars = reinterpret_cast<ConvBars*>(bias_s + 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = p.n_mtiles * p.n_ntiles;
    const int n_taps = p.tap_hi - p.tap_lo + 1;
    const int k_iters = n_taps * p.kb_per_tap;
    // Weights resident: when all K blocks of one N tile fit into the B ring's space (<= 128 KB: the three
    // 128-channel layers, 96 KB) and every tile of this CTA has the same N tile, B is loaded ONCE and only
    // A streams -- half the TMA / L2 traffic of those layers.
    const bool b_resident = k_iters * b_bytes <= CONV_STAGES * CONV_B_BYTES_MAX && gridDim.x % p.n_ntiles == 0;

    for (int i = threadIdx.x; i < p.c_out; i += blockDim.x) bias_s[i] = bias[i];
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < CONV_STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->tfull[a], 1);
            mbar_init(&bars->tempty[a], CONV_EPI_WARPS);
        }
        mbar_init(&bars->bfull, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(&bars->tmem_base, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, bars->tmem_base, 0);     // provably warp-uniform

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            if (b_resident && static_cast<int>(blockIdx.x) < n_tiles) {
                const int n0 = (blockIdx.x % p.n_ntiles) * p.nt;
                mbar_arrive_expect_tx(&bars->bfull, static_cast<uint32_t>(k_iters * b_bytes));
                int k = 0;
                for (int tap = p.tap_lo; tap <= p.tap_hi; ++tap)
                    for (int kb = 0; kb < p.kb_per_tap; ++kb, ++k)
                        tma_load_2d(b_s + k * b_bytes, &tmB, &bars->bfull, tap * p.c_in + kb * 64, n0);
            }
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int mt = tile / p.n_ntiles, ntile = tile % p.n_ntiles;
                int b0, f0;
                if (p.ms >= 128) { b0 = mt / p.tps; f0 = (mt % p.tps) * p.bf; }
                else             { b0 = mt * p.bb;  f0 = 0; }
                const int n0 = ntile * p.nt;
                for (int tap = p.tap_lo; tap <= p.tap_hi; ++tap) {
                    for (int kb = 0; kb < p.kb_per_tap; ++kb, ++it) {
                        const int s = it % CONV_STAGES;
                        const uint32_t ph = (it / CONV_STAGES) & 1;
                        mbar_wait(&bars->empty[s], ph ^ 1);
                        mbar_arrive_expect_tx(&bars->full[s], CONV_A_BYTES + (b_resident ? 0 : b_bytes));
                        uint8_t* da = a_s + s * CONV_A_BYTES;
                        const int c0 = kb * 64;
                        // mode 0: (c,t,f,b)        time conv, unit stride in the box: t = tap - pad
                        // mode 1: (c,tp,th,f,b)    time conv stride 2: t = 2 t' + tap -> parity tap&1, half t' + (tap>>1)
                        // mode 2: (c,t,fp,fh,b)    freq conv stride 2: f = 2 f' + tap
                        // mode 3: (c,t,f,b)        freq conv with a single output row: f = tap - pad
                        if (p.mode == 0)      tma_load_4d(da, &tmA, &bars->full[s], c0, tap - p.pad_lo, f0, b0);
                        else if (p.mode == 1) tma_load_5d(da, &tmA, &bars->full[s], c0, tap & 1, tap >> 1, f0, b0);
                        else if (p.mode == 2) tma_load_5d(da, &tmA, &bars->full[s], c0, 0, tap & 1, f0 + (tap >> 1), b0);
                        else                  tma_load_4d(da, &tmA, &bars->full[s], c0, 0, f0 + tap - p.pad_lo, b0);
                        if (!b_resident) tma_load_2d(b_s + s * CONV_B_BYTES_MAX, &tmB, &bars->full[s], tap * p.c_in + c0, n0);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (uniform control flow: descriptors and counters stay in uniform
        // registers, no ELECT / R2UR waterfall around every MMA); one elected lane issues.
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_f16(0u, 128, static_cast<uint32_t>(p.nt));
        const uint64_t adesc0 = umma_desc_sw128(smem_u32(a_s));
        const uint64_t bdesc0 = umma_desc_sw128(smem_u32(b_s));
        const uint32_t b_stride16 = static_cast<uint32_t>(b_resident ? b_bytes : CONV_B_BYTES_MAX) >> 4;
        if (b_resident && static_cast<int>(blockIdx.x) < n_tiles) mbar_wait(&bars->bfull, 0);
        int it = 0, lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            mbar_wait(&bars->tempty[acc], aph ^ 1);
            const uint32_t d_tmem = tmem_base + acc * 256;
            for (int k = 0; k < k_iters; ++k, ++it) {
                const int s = it % CONV_STAGES;
                const uint32_t ph = (it / CONV_STAGES) & 1;
                mbar_wait(&bars->full[s], ph);
                tc_fence_after();
                if (leader) {
                    const uint64_t adesc = adesc0 + static_cast<uint64_t>((s * CONV_A_BYTES) >> 4);
                    const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((b_resident ? k : s) * b_stride16);
#pragma unroll
                    for (int j = 0; j < 4; ++j) tc_mma_f16(d_tmem, adesc + 2 * j, bdesc + 2 * j, idesc, (k | j) != 0 ? 1u : 0u);
                    tc_commit(&bars->empty[s]);
                    if (k == k_iters - 1) tc_commit(&bars->tfull[acc]);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (CONV_EPI_WARPS warps)
        const int qd = warp & 3;               // TMEM lane quadrant = rows 32 qd .. 32 qd + 31 of the tile
        const int half = (warp - 2) >> 2;      // which part of the tile's columns
        const int cols = p.nt / CONV_EPI_PARTS;
        const int seg_len = p.ms < 32 ? p.ms : 32;     // rows of one segment inside this warp (power of two)
        int lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const int mt = tile / p.n_ntiles, ntile = tile % p.n_ntiles;
            const int acc = lt & 1;
            const uint32_t aph = (lt >> 1) & 1;
            const int m = mt * 128 + qd * 32 + lane;           // output row (NHWC position index)
            const bool row_ok = m < p.m_total;
            const int n_base = ntile * p.nt + half * cols;
            __half* yrow = y + static_cast<int64_t>(m) * p.c_out + n_base;
            mbar_wait(&bars->tfull[acc], aph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + acc * 256 + half * cols;
            float s1 = 0.f, s2 = 0.f;
            for (int c0 = 0; c0 < cols; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c0, v);
                tc_wait_ld();
                uint32_t pk[16];
                float ev[32];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float a = elu(__uint_as_float(v[j]) + bias_s[n_base + c0 + j]);
                    const float b = elu(__uint_as_float(v[j + 1]) + bias_s[n_base + c0 + j + 1]);
                    s1 += a + b;
                    s2 += a * a + b * b;
                    __half2 h = __floats2half2_rn(a, b);
                    pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
                    ev[j] = a;
                    ev[j + 1] = b;
                }
                if (row_ok && !p.y32) {
                    uint4* dst = reinterpret_cast<uint4*>(yrow + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                }
                if (row_ok && p.y32) {        // same element offsets, 4-byte elements
                    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + static_cast<int64_t>(m) * p.c_out + n_base + c0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(ev[4 * j], ev[4 * j + 1], ev[4 * j + 2], ev[4 * j + 3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tempty[acc]);
            // LayerNorm statistics: reduce over the rows of the same segment inside the warp, then one
            // slot per (segment, 32-row group, N tile, column half) -- no atomics, so the sums (and with
            // them every fingerprint) are bit-reproducible
            if (!row_ok) { s1 = 0.f; s2 = 0.f; }
            for (int o = 1; o < seg_len; o <<= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (row_ok && (lane & (seg_len - 1)) == 0) {
                const int b = m / p.ms;
                const int grp = p.ms >= 32 ? (m % p.ms) >> 5 : 0;
                const int64_t slot = ((static_cast<int64_t>(b) * p.groups + grp) * p.n_ntiles + ntile) * CONV_EPI_PARTS + half;
                reinterpret_cast<float2*>(part)[slot] = make_float2(s1, s2);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// LayerNorm over (F,T,C) with per-element gamma/beta: 8 elements per thread
// ------------------------------------------------------------------------------------------
constexpr int LN_SEGS = 8;         // segments per thread: gamma/beta (8 B per element) are loaded once for all of them
static_assert(LN_SEGS == 8, "ln_apply_kernel maps its 8 warps to the block's segments");
// The block first reduces the (sum, sum of squares) partial slots of its 8 segments -- one warp per segment,
// fixed order, double accumulation: bit-reproducible and no separate statistics launch.
__global__ void __launch_bounds__(256)
ln_apply_kernel(const __half* __restrict__ y, const float* __restrict__ part, int slots, float* __restrict__ stats_out,
                const float* __restrict__ gamma, const float* __restrict__ beta, __half* __restrict__ x, int per_seg, int n_seg,
                int c_out, int osplit, int y32) {
    __shared__ float2 mr_s[LN_SEGS];          // (mean, rstd) of the block's segments
    const int seg0 = blockIdx.y * LN_SEGS;
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;      // 8 warps = LN_SEGS segments
        const int seg = seg0 + w;
        if (seg < n_seg) {
            const float2* p = reinterpret_cast<const float2*>(part) + static_cast<int64_t>(seg) * slots;
            double a = 0.0, b = 0.0;
            for (int i = lane; i < slots; i += 32) {
                const float2 v = p[i];
                a += v.x;
                b += v.y;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                b += __shfl_xor_sync(0xffffffffu, b, o);
            }
            if (lane == 0) {
                const float inv_n = 1.f / static_cast<float>(per_seg);
                const float s1 = static_cast<float>(a), s2 = static_cast<float>(b);
                const float mean = s1 * inv_n;
                const float var = fmaxf(s2 * inv_n - mean * mean, 0.f);
                mr_s[w] = make_float2(mean, rsqrtf(var + LN_EPS));
                if (blockIdx.x == 0 && stats_out) {
                    stats_out[2 * seg] = s1;
                    stats_out[2 * seg + 1] = s2;
                }
            }
        }
    }
    __syncthreads();
    const int off = (blockIdx.x * blockDim.x + threadIdx.x) * 8;        // 8 consecutive elements of the (F,T,C) volume
    if (off >= per_seg) return;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + off), g1 = *reinterpret_cast<const float4*>(gamma + off + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + off), b1 = *reinterpret_cast<const float4*>(beta + off + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll 4
    for (int k = 0; k < LN_SEGS; ++k) {
        const int b = seg0 + k;
        if (b >= n_seg) break;
        const float mean = mr_s[k].x, rstd = mr_s[k].y;
        const int64_t idx = static_cast<int64_t>(b) * per_seg + off;
        float fv[8];
        if (y32) {
            const float4 r0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(y) + idx);
            const float4 r1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(y) + idx + 4);
            fv[0] = r0.x; fv[1] = r0.y; fv[2] = r0.z; fv[3] = r0.w; fv[4] = r1.x; fv[5] = r1.y; fv[6] = r1.z; fv[7] = r1.w;
        } else {
            const uint4 raw = *reinterpret_cast<const uint4*>(y + idx);
            const __half2* hv = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(hv[j]);
                fv[2 * j] = f.x;
                fv[2 * j + 1] = f.y;
            }
        }
        uint4 outv, lowv;
        uint32_t* ov = reinterpret_cast<uint32_t*>(&outv);
        uint32_t* lv = reinterpret_cast<uint32_t*>(&lowv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = (fv[2 * j] - mean) * rstd * gg[2 * j] + bb[2 * j];
            const float c = (fv[2 * j + 1] - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1];
            __half2 h = __floats2half2_rn(a, c);
            ov[j] = *reinterpret_cast<uint32_t*>(&h);
            const float2 back = __half22float2(h);
            __half2 lo = __floats2half2_rn(a - back.x, c - back.y);
            lv[j] = *reinterpret_cast<uint32_t*>(&lo);
        }
        if (osplit == 1) {
            *reinterpret_cast<uint4*>(x + idx) = outv;
        } else {            // [hi | hi | lo] per position
            const int pos = off / c_out, ch = off - pos * c_out;
            __half* dst = x + (static_cast<int64_t>(b) * per_seg + static_cast<int64_t>(pos) * c_out) * 3 + ch;
            *reinterpret_cast<uint4*>(dst) = outv;
            *reinterpret_cast<uint4*>(dst + c_out) = outv;
            *reinterpret_cast<uint4*>(dst + 2 * c_out) = lowv;
        }
    }
}

// ------------------------------------------------------------------------------------------
// divide-and-encode head (nnfp.py:132-156): block (q, 128 segments) -- the 8x32 + 32 weights of slice q
// sit in shared memory and are broadcast to the 128 segments; then one warp per segment L2-normalises.
// x: (B, 1024) normalised fp16 (Flatten of (1,1,1024)); slice q = features 8q .. 8q+7 (nnfp.py:155)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
divenc_kernel(const __half* __restrict__ x, int n_seg, const float* __restrict__ w1, const float* __restrict__ b1,
              const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ raw) {
    __shared__ float W1[8 * 32], B1[32], W2[32];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < 256; i += 128) W1[i] = w1[q * 256 + i];
    if (threadIdx.x < 32) {
        B1[threadIdx.x] = b1[q * 32 + threadIdx.x];
        W2[threadIdx.x] = w2[q * 32 + threadIdx.x];
    }
    __syncthreads();
    const int seg = blockIdx.y * 128 + threadIdx.x;
    if (seg >= n_seg) return;
    // the last activation is stored [hi | hi | lo] (1024 channels each): hi + lo is the fp32 value to 2^-22
    const uint4 rv = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(seg) * 3072 + q * 8);
    const uint4 rl = *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(seg) * 3072 + 2048 + q * 8);
    const __half2* hv = reinterpret_cast<const __half2*>(&rv);
    const __half2* hl = reinterpret_cast<const __half2*>(&rl);
    float in[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(hv[j]), g = __half22float2(hl[j]);
        in[2 * j] = f.x + g.x;
        in[2 * j + 1] = f.y + g.y;
    }
    float acc = b2[q];
#pragma unroll 8
    for (int u = 0; u < 32; ++u) {
        float hsum = B1[u];
#pragma unroll
        for (int sdim = 0; sdim < 8; ++sdim) hsum += in[sdim] * W1[sdim * 32 + u];
        hsum = hsum > 0.f ? hsum : expm1f(hsum);
        acc += hsum * W2[u];
    }
    raw[static_cast<int64_t>(seg) * EMB + q] = acc;
}

__global__ void __launch_bounds__(256)
l2norm_kernel(const float* __restrict__ raw, int n_seg, float* __restrict__ emb) {
    const int lane = threadIdx.x & 31;
    const int seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (seg >= n_seg) return;
    const float4 v = reinterpret_cast<const float4*>(raw + static_cast<int64_t>(seg) * EMB)[lane];
    float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(fmaxf(ss, L2_EPS));
    reinterpret_cast<float4*>(emb + static_cast<int64_t>(seg) * EMB)[lane] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
}

__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = static_cast<float>(in[i]) * (1.0f / 32768.0f);
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static int encoder_init(nafp_ctx* ctx) {
    if (ctx->encoder) return NAFP_OK;
    EncoderState* s = new EncoderState();
    build_geometry(s->g);
    for (int l = 0; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        const size_t per = static_cast<size_t>(L.ms) * L.c_out;
        NAFP_CUDA(cudaMalloc(&s->bias[l], L.c_out * sizeof(float)));
        NAFP_CUDA(cudaMalloc(&s->ln_g[l], per * sizeof(float)));
        NAFP_CUDA(cudaMalloc(&s->ln_b[l], per * sizeof(float)));
        if (l == 0) NAFP_CUDA(cudaMalloc(&s->w0, 3 * 128 * sizeof(float)));
        else NAFP_CUDA(cudaMalloc(&s->wt[l], static_cast<size_t>(L.c_out) * 3 * L.c_in * L.ksplit * sizeof(__half)));
    }
    NAFP_CUDA(cudaMalloc(&s->dw1, 128 * 8 * 32 * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->db1, 128 * 32 * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->dw2, 128 * 32 * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->db2, 128 * sizeof(float)));
    for (int l = 1; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        const uint64_t C = static_cast<uint64_t>(L.c_in) * L.ksplit;
        const uint64_t wd[2] = {3 * C, static_cast<uint64_t>(L.c_out)};
        const uint64_t ws[2] = {2, 3 * C * 2};
        const uint32_t wb[2] = {64, static_cast<uint32_t>(L.nt)};
        NAFP_TRY(make_tensor_map(&s->tmB[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, s->wt[l], wd, ws, wb, nullptr,
                                 CU_TENSOR_MAP_SWIZZLE_128B));
    }
    NAFP_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    NAFP_CUDA(cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        NAFP_CUDA(cudaEventCreateWithFlags(&s->ev_up[b], cudaEventDisableTiming));
        NAFP_CUDA(cudaEventCreateWithFlags(&s->ev_free[b], cudaEventDisableTiming));
    }
    ctx->encoder = s;
    return NAFP_OK;
}

static void encoder_release_arena(EncoderState* s) {
    for (int l = 0; l < ENC_LAYERS; ++l) {
        if (s->x[l]) cudaFree(s->x[l]);
        s->x[l] = nullptr;
    }
    void** bufs[] = {reinterpret_cast<void**>(&s->y), reinterpret_cast<void**>(&s->stats), reinterpret_cast<void**>(&s->part),
                     reinterpret_cast<void**>(&s->raw), reinterpret_cast<void**>(&s->mel), &s->xin[0], &s->xin[1],
                     reinterpret_cast<void**>(&s->emb)};
    for (void** b : bufs) {
        if (*b) cudaFree(*b);
        *b = nullptr;
    }
    s->cap = 0;
    s->last_n = 0;
}

// The activation arena (2.3 MB per segment) and the activation tensor maps are sized for `cap` segments; a pass of
// more segments than that re-allocates them (ENC_CHUNK_MIN, then the request rounded up to 1,000, <= ENC_CHUNK_MAX).
static int encoder_reserve(nafp_ctx* ctx, int64_t n) {
    EncoderState* s = ctx->encoder;
    if (n <= s->cap) return NAFP_OK;
    int64_t cap = (n + 999) / 1000 * 1000;
    if (cap < ENC_CHUNK_MIN) cap = ENC_CHUNK_MIN;
    if (cap > ENC_CHUNK_MAX) cap = ENC_CHUNK_MAX;
    NAFP_REQUIRE(n <= cap, NAFP_ERR_INVALID, "encoder: %lld segments in one pass (at most %d)", (long long)n, ENC_CHUNK_MAX);
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    encoder_release_arena(s);
    size_t ymax = 0;
    for (int l = 0; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        const size_t per = static_cast<size_t>(L.ms) * L.c_out;
        if (per > ymax) ymax = per;
        // +128 rows of slack: the GEMM epilogue never writes past m_total, TMA boxes may read past it
        const size_t x_elems = (static_cast<size_t>(cap) * per + 128 * L.c_out) * L.osplit;
        NAFP_CUDA(cudaMalloc(&s->x[l], x_elems * sizeof(__half)));
        NAFP_CUDA(cudaMemset(s->x[l], 0, x_elems * sizeof(__half)));
    }
    NAFP_CUDA(cudaMalloc(&s->y, (static_cast<size_t>(cap) * ymax + 128 * 1024) * sizeof(__half)));
    NAFP_CUDA(cudaMalloc(&s->stats, static_cast<size_t>(ENC_LAYERS) * cap * 2 * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->part, static_cast<size_t>(cap) * 64 * CONV_EPI_PARTS * 2 * sizeof(float)));   // <= 64 row groups x parts float2 slots per segment
    NAFP_CUDA(cudaMalloc(&s->mel, static_cast<size_t>(cap) * 8192 * sizeof(float)));
    for (int b = 0; b < 2; ++b) NAFP_CUDA(cudaMalloc(&s->xin[b], static_cast<size_t>(cap) * 8000 * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->emb, static_cast<size_t>(cap) * EMB * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->raw, static_cast<size_t>(cap) * EMB * sizeof(float)));
    // activation tensor maps (rows past the live batch are masked by the epilogue)
    for (int l = 1; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        const uint64_t C = static_cast<uint64_t>(L.c_in) * L.ksplit, T = L.t_in, F = L.f_in, B = cap;
        uint64_t dims[5], str[5];
        uint32_t box[5];
        int rank;
        if (L.mode == 0 || L.mode == 3) {
            rank = 4;
            dims[0] = C; dims[1] = T; dims[2] = F; dims[3] = B;
            str[0] = 2; str[1] = C * 2; str[2] = T * C * 2; str[3] = F * T * C * 2;
            box[0] = 64; box[1] = L.bt; box[2] = L.bf; box[3] = L.bb;
        } else if (L.mode == 1) {
            rank = 5;
            dims[0] = C; dims[1] = 2; dims[2] = T / 2; dims[3] = F; dims[4] = B;
            str[0] = 2; str[1] = C * 2; str[2] = 2 * C * 2; str[3] = T * C * 2; str[4] = F * T * C * 2;
            box[0] = 64; box[1] = 1; box[2] = L.bt; box[3] = L.bf; box[4] = L.bb;
        } else {
            rank = 5;
            dims[0] = C; dims[1] = T; dims[2] = 2; dims[3] = F / 2; dims[4] = B;
            str[0] = 2; str[1] = C * 2; str[2] = T * C * 2; str[3] = 2 * T * C * 2; str[4] = F * T * C * 2;
            box[0] = 64; box[1] = L.bt; box[2] = 1; box[3] = L.bf; box[4] = L.bb;
        }
        NAFP_TRY(make_tensor_map(&s->tmA[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, s->x[l - 1], dims, str, box, nullptr,
                                 CU_TENSOR_MAP_SWIZZLE_128B));
    }
    s->cap = static_cast<int>(cap);
    return NAFP_OK;
}

void encoder_destroy(nafp_ctx* ctx) {
    EncoderState* s = ctx->encoder;
    if (!s) return;
    encoder_release_arena(s);
    for (int l = 0; l < ENC_LAYERS; ++l) {
        cudaFree(s->bias[l]); cudaFree(s->ln_g[l]); cudaFree(s->ln_b[l]);
        if (s->wt[l]) cudaFree(s->wt[l]);
    }
    void* bufs[] = {s->w0, s->dw1, s->db1, s->dw2, s->db2};
    for (void* b : bufs) if (b) cudaFree(b);
    for (int b = 0; b < 2; ++b) {
        if (s->ev_up[b]) cudaEventDestroy(s->ev_up[b]);
        if (s->ev_free[b]) cudaEventDestroy(s->ev_free[b]);
    }
    if (s->copy_stream) cudaStreamDestroy(s->copy_stream);
    delete s;
    ctx->encoder = nullptr;
}

// one pass over n <= cap segments (encoder_reserve); mel is either final (gmax == nullptr) or raw log-mel + group maxima
static int encoder_pass(nafp_ctx* ctx, const float* mel, const int32_t* gmax, int64_t group_size, int64_t seg0, int n,
                        float* emb_dev) {
    EncoderState* s = ctx->encoder;
    cudaStream_t st = ctx->stream;
    for (int l = 0; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        float* stats = s->stats + static_cast<size_t>(l) * s->cap * 2;
        const int per = L.ms * L.c_out;
        if (l == 0) {
            conv0_stats_kernel<<<dim3(8, n), 256, 0, st>>>(mel, gmax, group_size, n, s->w0, s->bias[0], s->part);
            conv0_ln_kernel<<<dim3(8, (n + CONV0_SEGS - 1) / CONV0_SEGS), 256, 0, st>>>(
                mel, gmax, group_size, n, s->w0, s->bias[0], s->part, stats, s->ln_g[0], s->ln_b[0], s->x[0]);
            ctx->launches += 2;
            continue;
        }
        ConvParams p;
        p.m_total = n * L.ms; p.ms = L.ms; p.c_in = L.c_in * L.ksplit; p.c_out = L.c_out; p.nt = L.nt;
        p.n_ntiles = L.c_out / L.nt; p.n_mtiles = (p.m_total + 127) / 128;
        p.mode = L.mode; p.pad_lo = L.pad_lo; p.tap_lo = L.tap_lo; p.tap_hi = L.tap_hi;
        p.kb_per_tap = L.c_in * L.ksplit / 64; p.tps = L.ms >= 128 ? L.ms / 128 : 0; p.bf = L.bf; p.bb = L.bb;
        p.groups = L.ms >= 32 ? L.ms / 32 : 1;
        p.y32 = l >= ENC_Y32_FROM ? 1 : 0;      // small layers keep their pre-LN output in fp32 (one fp16 rounding less)
        const int slots = p.groups * p.n_ntiles * CONV_EPI_PARTS;
        const int tiles = p.n_mtiles * p.n_ntiles;
        const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
        conv_gemm_kernel<<<grid, CONV_THREADS, CONV_SMEM, st>>>(s->tmA[l], s->tmB[l], p, s->bias[l], s->y, s->part);
        ln_apply_kernel<<<dim3((per / 8 + 255) / 256, (n + LN_SEGS - 1) / LN_SEGS), 256, 0, st>>>(
            s->y, s->part, slots, stats, s->ln_g[l], s->ln_b[l], s->x[l], per, n, L.c_out, L.osplit, p.y32);
        ctx->launches += 2;
    }
    divenc_kernel<<<dim3(EMB, (n + 127) / 128), 128, 0, st>>>(s->x[ENC_LAYERS - 1], n, s->dw1, s->db1, s->dw2, s->db2, s->raw);
    l2norm_kernel<<<(n * 32 + 255) / 256, 256, 0, st>>>(s->raw, n, emb_dev);
    ctx->launches += 2;
    NAFP_CUDA(cudaGetLastError());
    s->last_n = n;
    return NAFP_OK;
}

static int fingerprint_dev(nafp_ctx* ctx, const void* x_dev, bool pcm16, int64_t n_seg, int64_t group_size,
                           float* emb_dev, const int64_t* seg_off = nullptr, const int32_t* seg_valid = nullptr) {
    EncoderState* s = ctx->encoder;
    // chunks are whole groups so that every group's batch-global max is complete (melspectrogram.py:108)
    int64_t chunk = group_size <= ENC_CHUNK_MAX ? (ENC_CHUNK_MAX / group_size) * group_size : 0;
    NAFP_REQUIRE(chunk > 0, NAFP_ERR_UNSUPPORTED, "fingerprint: group_size %lld exceeds the %d-segment encoder pass",
                 (long long)group_size, ENC_CHUNK_MAX);
    NAFP_TRY(encoder_reserve(ctx, n_seg < chunk ? n_seg : chunk));
    const size_t elt = pcm16 ? sizeof(int16_t) : sizeof(float);
    for (int64_t s0 = 0; s0 < n_seg; s0 += chunk) {
        const int n = static_cast<int>(n_seg - s0 < chunk ? n_seg - s0 : chunk);
        const int32_t* gmax = nullptr;
        // (n_seg, 8000) rows, or windows of track sample runs addressed through seg_off (absolute sample offsets)
        const void* xin = seg_off ? x_dev : static_cast<const uint8_t*>(x_dev) + static_cast<size_t>(s0) * 8000 * elt;
        NAFP_TRY(logmel_run(ctx, xin, pcm16, n, group_size, s->mel, false, &gmax, seg_off ? seg_off + s0 : nullptr,
                            seg_valid ? seg_valid + s0 : nullptr));
        NAFP_TRY(encoder_pass(ctx, s->mel, gmax, group_size, 0, n, emb_dev + s0 * EMB));
    }
    return NAFP_OK;
}

}  // namespace nafp

using namespace nafp;

extern "C" {

int nafp_weights_load(nafp_ctx* ctx, const float* const* conv_w, const float* const* conv_b, const float* const* ln_g,
                      const float* const* ln_b, const float* div_w1, const float* div_b1, const float* div_w2,
                      const float* div_b2) {
    NAFP_REQUIRE(ctx && conv_w && conv_b && ln_g && ln_b && div_w1 && div_b1 && div_w2 && div_b2, NAFP_ERR_INVALID,
                 "nafp_weights_load: NULL argument");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    NAFP_TRY(encoder_init(ctx));
    EncoderState* s = ctx->encoder;
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int l = 0; l < ENC_LAYERS; ++l) {
        const ConvGeom& L = s->g[l];
        NAFP_REQUIRE(conv_w[l] && conv_b[l] && ln_g[l] && ln_b[l], NAFP_ERR_INVALID, "nafp_weights_load: layer %d NULL", l);
        const size_t per = static_cast<size_t>(L.ms) * L.c_out;
        NAFP_CUDA(cudaMemcpy(s->bias[l], conv_b[l], L.c_out * sizeof(float), cudaMemcpyHostToDevice));
        NAFP_CUDA(cudaMemcpy(s->ln_g[l], ln_g[l], per * sizeof(float), cudaMemcpyHostToDevice));
        NAFP_CUDA(cudaMemcpy(s->ln_b[l], ln_b[l], per * sizeof(float), cudaMemcpyHostToDevice));
        if (l == 0) {
            NAFP_CUDA(cudaMemcpy(s->w0, conv_w[0], 3 * 128 * sizeof(float), cudaMemcpyHostToDevice));
        } else {
            // HWIO [tap][cin][cout] -> K-major B operand [cout][tap*cin + cin] in fp16
            // split layers: per tap [hi | lo | hi], matching the [hi | hi | lo] activations
            const int ce = L.c_in * L.ksplit, K = 3 * ce;
            std::vector<__half> wt(static_cast<size_t>(L.c_out) * K);
            for (int tap = 0; tap < 3; ++tap)
                for (int ci = 0; ci < L.c_in; ++ci) {
                    const float* src = conv_w[l] + (static_cast<size_t>(tap) * L.c_in + ci) * L.c_out;
                    for (int co = 0; co < L.c_out; ++co) {
                        const __half hi = __float2half_rn(src[co]);
                        __half* row = &wt[static_cast<size_t>(co) * K + tap * ce];
                        row[ci] = hi;
                        if (L.ksplit == 3) {
                            row[L.c_in + ci] = __float2half_rn(src[co] - __half2float(hi));
                            row[2 * L.c_in + ci] = hi;
                        }
                    }
                }
            NAFP_CUDA(cudaMemcpy(s->wt[l], wt.data(), wt.size() * sizeof(__half), cudaMemcpyHostToDevice));
        }
    }
    NAFP_CUDA(cudaMemcpy(s->dw1, div_w1, 128 * 8 * 32 * sizeof(float), cudaMemcpyHostToDevice));
    NAFP_CUDA(cudaMemcpy(s->db1, div_b1, 128 * 32 * sizeof(float), cudaMemcpyHostToDevice));
    NAFP_CUDA(cudaMemcpy(s->dw2, div_w2, 128 * 32 * sizeof(float), cudaMemcpyHostToDevice));
    NAFP_CUDA(cudaMemcpy(s->db2, div_b2, 128 * sizeof(float), cudaMemcpyHostToDevice));
    s->weights = true;
    return NAFP_OK;
}

int nafp_encoder_forward(nafp_ctx* ctx, const float* mel_dev, int64_t n_seg, float* emb_dev) {
    NAFP_REQUIRE(ctx && n_seg >= 0 && (n_seg == 0 || (mel_dev && emb_dev)), NAFP_ERR_INVALID,
                 "nafp_encoder_forward: bad arguments");
    NAFP_REQUIRE(ctx->encoder && ctx->encoder->weights, NAFP_ERR_STATE, "nafp_encoder_forward: call nafp_weights_load first");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    NAFP_TRY(encoder_reserve(ctx, n_seg < ENC_CHUNK_MAX ? n_seg : ENC_CHUNK_MAX));
    for (int64_t s0 = 0; s0 < n_seg; s0 += ENC_CHUNK_MAX) {
        const int n = static_cast<int>(n_seg - s0 < ENC_CHUNK_MAX ? n_seg - s0 : ENC_CHUNK_MAX);
        NAFP_TRY(encoder_pass(ctx, mel_dev + s0 * 8192, nullptr, 1, s0, n, emb_dev + s0 * EMB));
    }
    return NAFP_OK;
}

int nafp_fingerprint(nafp_ctx* ctx, const float* x_dev, int64_t n_seg, int64_t group_size, float* emb_dev) {
    NAFP_REQUIRE(ctx && n_seg >= 0 && group_size >= 1 && (n_seg == 0 || (x_dev && emb_dev)), NAFP_ERR_INVALID,
                 "nafp_fingerprint: bad arguments");
    NAFP_REQUIRE(ctx->encoder && ctx->encoder->weights, NAFP_ERR_STATE, "nafp_fingerprint: call nafp_weights_load first");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    return fingerprint_dev(ctx, x_dev, false, n_seg, group_size, emb_dev);
}

static int fingerprint_host(nafp_ctx* ctx, const void* x_host, bool pcm16, int64_t n_seg, int64_t group_size,
                            float* emb_host) {
    NAFP_REQUIRE(ctx && n_seg >= 0 && group_size >= 1 && (n_seg == 0 || (x_host && emb_host)), NAFP_ERR_INVALID,
                 "nafp_fingerprint_host: bad arguments");
    NAFP_REQUIRE(ctx->encoder && ctx->encoder->weights, NAFP_ERR_STATE, "nafp_fingerprint_host: call nafp_weights_load first");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    EncoderState* s = ctx->encoder;
    const int64_t chunk = group_size <= ENC_CHUNK_MAX ? (ENC_CHUNK_MAX / group_size) * group_size : 0;
    NAFP_REQUIRE(chunk > 0, NAFP_ERR_UNSUPPORTED, "fingerprint: group_size %lld exceeds the %d-segment encoder pass",
                 (long long)group_size, ENC_CHUNK_MAX);
    NAFP_TRY(encoder_reserve(ctx, n_seg < chunk ? n_seg : chunk));      // before s->xin / s->emb are read below
    const size_t elt = pcm16 ? sizeof(int16_t) : sizeof(float);
    const uint8_t* src = static_cast<const uint8_t*>(x_host);
    auto upload = [&](int64_t s0, int b) -> cudaError_t {      // pass starting at s0 -> xin[b], on the copy stream
        const int64_t n = n_seg - s0 < chunk ? n_seg - s0 : chunk;
        cudaError_t e = cudaStreamWaitEvent(s->copy_stream, s->ev_free[b], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(s->xin[b], src + static_cast<size_t>(s0) * 8000 * elt, static_cast<size_t>(n) * 8000 * elt,
                                cudaMemcpyHostToDevice, s->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(s->ev_up[b], s->copy_stream);
        return e;
    };
    // the copy stream starts behind everything already queued on the compute stream (it may still read xin[0/1])
    for (int b = 0; b < 2; ++b) NAFP_CUDA(cudaEventRecord(s->ev_free[b], ctx->stream));
    if (n_seg > 0) NAFP_CUDA(upload(0, 0));
    int b = 0;
    for (int64_t s0 = 0; s0 < n_seg; s0 += chunk, b ^= 1) {
        const int64_t n = n_seg - s0 < chunk ? n_seg - s0 : chunk;
        if (s0 + chunk < n_seg) NAFP_CUDA(upload(s0 + chunk, b ^ 1));     // under this pass's kernels
        NAFP_CUDA(cudaStreamWaitEvent(ctx->stream, s->ev_up[b], 0));
        NAFP_TRY(fingerprint_dev(ctx, s->xin[b], pcm16, n, group_size, s->emb));
        NAFP_CUDA(cudaEventRecord(s->ev_free[b], ctx->stream));
        NAFP_CUDA(cudaMemcpyAsync(emb_host + s0 * EMB, s->emb, static_cast<size_t>(n) * EMB * sizeof(float),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    }
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    return NAFP_OK;
}

int nafp_fingerprint_host(nafp_ctx* ctx, const float* x_host, int64_t n_seg, int64_t group_size, float* emb_host) {
    return fingerprint_host(ctx, x_host, false, n_seg, group_size, emb_host);
}
int nafp_fingerprint_pcm16_host(nafp_ctx* ctx, const int16_t* pcm_host, int64_t n_seg, int64_t group_size,
                                float* emb_host) {
    return fingerprint_host(ctx, pcm_host, true, n_seg, group_size, emb_host);
}

// generate.py's path: whole-track int16 sample runs + one (offset, valid length) pair per segment; the overlapping
// 1 s segments (0.5 s hop) are cut by the log-mel kernel, so every sample is uploaded once instead of twice and the
// host never materialises the (n_seg, 8000) array.
int nafp_fingerprint_pcm16_tracks_host(nafp_ctx* ctx, const int16_t* pcm_host, int64_t n_samples, const int64_t* seg_off_host,
                                       const int32_t* seg_valid_host, int64_t n_seg, int64_t group_size, float* emb_host) {
    NAFP_REQUIRE(ctx && n_seg >= 0 && n_samples >= 0 && group_size >= 1 &&
                     (n_seg == 0 || (pcm_host && seg_off_host && seg_valid_host && emb_host)),
                 NAFP_ERR_INVALID, "nafp_fingerprint_pcm16_tracks_host: bad arguments");
    NAFP_REQUIRE(ctx->encoder && ctx->encoder->weights, NAFP_ERR_STATE,
                 "nafp_fingerprint_pcm16_tracks_host: call nafp_weights_load first");
    if (n_seg == 0) return NAFP_OK;
    for (int64_t i = 0; i < n_seg; ++i)
        NAFP_REQUIRE(seg_off_host[i] >= 0 && seg_valid_host[i] >= 0 && seg_valid_host[i] <= 8000 &&
                         seg_off_host[i] + seg_valid_host[i] <= n_samples,
                     NAFP_ERR_INVALID, "nafp_fingerprint_pcm16_tracks_host: segment %lld reads [%lld, +%d) of %lld samples",
                     (long long)i, (long long)seg_off_host[i], (int)seg_valid_host[i], (long long)n_samples);
    NAFP_CUDA(cudaSetDevice(ctx->device));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));      // the staging buffer may be re-allocated below
    // one staging buffer, every part 256-byte aligned: samples (+1 so that it is never empty), offsets, valid
    // lengths, fingerprints
    const int64_t pcm_bytes = (n_samples + 1) * 2, pcm_pad = (pcm_bytes + 255) / 256 * 256;
    const int64_t off_bytes = (n_seg * 8 + 255) / 256 * 256, val_bytes = (n_seg * 4 + 255) / 256 * 256;
    const int64_t emb_bytes = n_seg * EMB * static_cast<int64_t>(sizeof(float));
    NAFP_TRY(ensure_dev(ctx, &ctx->stage_dev, &ctx->stage_dev_bytes, pcm_pad + off_bytes + val_bytes + emb_bytes));
    uint8_t* base = static_cast<uint8_t*>(ctx->stage_dev);
    int16_t* pcm_dev = reinterpret_cast<int16_t*>(base);
    int64_t* off_dev = reinterpret_cast<int64_t*>(base + pcm_pad);
    int32_t* val_dev = reinterpret_cast<int32_t*>(base + pcm_pad + off_bytes);
    float* emb_dev = reinterpret_cast<float*>(base + pcm_pad + off_bytes + val_bytes);
    NAFP_CUDA(cudaMemcpyAsync(pcm_dev, pcm_host, static_cast<size_t>(n_samples) * 2, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_CUDA(cudaMemcpyAsync(off_dev, seg_off_host, static_cast<size_t>(n_seg) * 8, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_CUDA(cudaMemcpyAsync(val_dev, seg_valid_host, static_cast<size_t>(n_seg) * 4, cudaMemcpyHostToDevice, ctx->stream));
    NAFP_TRY(fingerprint_dev(ctx, pcm_dev, true, n_seg, group_size, emb_dev, off_dev, val_dev));
    NAFP_CUDA(cudaMemcpyAsync(emb_host, emb_dev, static_cast<size_t>(emb_bytes), cudaMemcpyDeviceToHost, ctx->stream));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    return NAFP_OK;
}

int nafp_encoder_activation_host(nafp_ctx* ctx, int layer, int64_t n_seg, float* out_host) {
    NAFP_REQUIRE(ctx && ctx->encoder && out_host && layer >= 0 && layer < ENC_LAYERS, NAFP_ERR_INVALID,
                 "nafp_encoder_activation_host: bad arguments");
    EncoderState* s = ctx->encoder;
    NAFP_REQUIRE(n_seg >= 0 && n_seg <= s->last_n, NAFP_ERR_INVALID,
                 "nafp_encoder_activation_host: %lld segments requested, last pass had %lld", (long long)n_seg,
                 (long long)s->last_n);
    const ConvGeom& L = s->g[layer];
    const size_t n = static_cast<size_t>(n_seg) * L.ms * L.c_out;
    std::vector<__half> tmp(n * L.osplit);
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    NAFP_CUDA(cudaMemcpy(tmp.data(), s->x[layer], tmp.size() * sizeof(__half), cudaMemcpyDeviceToHost));
    if (L.osplit == 1) {
        for (size_t i = 0; i < n; ++i) out_host[i] = __half2float(tmp[i]);
    } else {                // [hi | hi | lo] per position
        const size_t C = L.c_out;
        for (size_t i = 0; i < n; ++i) {
            const size_t pos = i / C, c = i % C;
            out_host[i] = __half2float(tmp[pos * 3 * C + c]) + __half2float(tmp[pos * 3 * C + 2 * C + c]);
        }
    }
    return NAFP_OK;
}

}  // extern "C"
