// Fused log-mel front end: replaces the kapre STFT / Magnitude / ApplyFilterbank
// chain + the log / max / clamp of Melspec_layer.call (model/fp/melspec/melspectrogram.py:59-112).
//
// One CTA per 1 s segment.  The 8000 samples are staged once in shared memory (zero padded to
// 9024, :59-65); four 64-thread groups each take one of the 32 frames at a time: periodic Hann
// window, 1024-point real FFT as a 512-point complex FFT (three radix-8 passes, registers +
// shared-memory transposes) and a split step, |X| for bins 39..511 (the only bins with a non-zero
// mel weight), the sparse 256-band Slaney filterbank (2..8 taps per band), +0.06, log10.  The
// (256, 32) tile leaves the CTA once, already in the (F, T) order of the reference's Permute, and
// the per-group maximum (the batch-global max of :108) is folded in with an atomic; the subtraction
// and the -80 clamp are a second elementwise pass (or are applied by the encoder's first layer).
#include <cmath>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace nafp {

constexpr int SEG_LEN = 8000;
constexpr int NFFT = 1024;
constexpr int HOP = 256;
constexpr int PAD = NFFT / 2;
constexpr int PADDED = SEG_LEN + 2 * PAD;            // 9024
constexpr int NFRAMES = 1 + (PADDED - NFFT) / HOP;   // 32
constexpr int NMEL = 256;
constexpr int MAXTAPS = 8;
constexpr int TILE_LD = NFRAMES + 1;                 // 33
constexpr int S_LD = 68;                             // stage-1 exchange row stride (floats)
constexpr int GRP_FLOATS = 2 * 8 * S_LD + 2 * 64 * 9;   // P (S / Z) + Q (S2 / mag) per group
constexpr int LOGMEL_SMEM = (PADDED + 4 * GRP_FLOATS + NMEL * TILE_LD) * 4;
static_assert(2 * 8 * S_LD >= 1024, "P must hold Z");
static_assert(2 * 64 * 9 >= 512, "Q must hold mag");

struct LogmelState {
    float* melw = nullptr;     // [MAXTAPS][NMEL]
    int* start = nullptr;      // [NMEL]
    float* tab = nullptr;      // window + twiddle tables (TAB_* offsets), computed in double on the host
    int32_t* gmax = nullptr;   // [2][cap] ordered-int group maxima, then group minima
    int64_t gmax_cap = 0;
    bool segment_norm = false; // MODEL.FEAT == 'melspec_maxnorm' (melspectrogram.py:110-111)
};

// Window and twiddle factors come from tables (computed once in double precision): every thread needs 16 window values
// and 24 twiddles, and 40 sincospif calls per thread were a fifth of the kernel's instructions.
constexpr int TAB_WIN = 0;                  // [1024]        periodic Hann window
constexpr int TAB_TW1 = 1024;               // [512][2]      e^{-2 pi i k / 512}
constexpr int TAB_TW2 = TAB_TW1 + 2 * 512;  // [64][2]       e^{-2 pi i k / 64}
constexpr int TAB_TWP = TAB_TW2 + 2 * 64;   // [512][2]      e^{-2 pi i k / 1024}
constexpr int TAB_FLOATS = TAB_TWP + 2 * 512;

// ---- 8-point FFT, natural order in and out
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // * (-i)

__device__ __forceinline__ void fft4(float2& c0, float2& c1, float2& c2, float2& c3) {
    const float2 d0 = cadd(c0, c2), d2 = csub(c0, c2), d1 = cadd(c1, c3), d3 = mul_mi(csub(c1, c3));
    c0 = cadd(d0, d1);
    c2 = csub(d0, d1);
    c1 = cadd(d2, d3);
    c3 = csub(d2, d3);
}
__device__ __forceinline__ void fft8(float2 (&a)[8]) {
    const float r = 0.70710678118654752f;
    float2 b0 = cadd(a[0], a[4]), b4 = csub(a[0], a[4]);
    float2 b1 = cadd(a[1], a[5]), b5 = csub(a[1], a[5]);
    float2 b2 = cadd(a[2], a[6]), b6 = csub(a[2], a[6]);
    float2 b3 = cadd(a[3], a[7]), b7 = csub(a[3], a[7]);
    b5 = make_float2((b5.x + b5.y) * r, (b5.y - b5.x) * r);      // * W8^1
    b6 = mul_mi(b6);                                             // * W8^2
    b7 = make_float2((b7.y - b7.x) * r, -(b7.x + b7.y) * r);     // * W8^3
    fft4(b0, b1, b2, b3);
    fft4(b4, b5, b6, b7);
    a[0] = b0; a[2] = b1; a[4] = b2; a[6] = b3;
    a[1] = b4; a[3] = b5; a[5] = b6; a[7] = b7;
}
template <typename TIn>
__device__ __forceinline__ float to_sample(TIn v);
template <>
__device__ __forceinline__ float to_sample<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_sample<int16_t>(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }

template <typename TIn>
__global__ void __launch_bounds__(256, 2)
logmel_kernel(const TIn* __restrict__ x, int64_t n_seg, int64_t group_size, const float* __restrict__ melw,
              const int* __restrict__ start, const float* __restrict__ tab, float* __restrict__ out, int32_t* __restrict__ gmax,
              int32_t* __restrict__ gmin, const int64_t* __restrict__ seg_off, const int32_t* __restrict__ seg_valid) {
    extern __shared__ float smem[];
    float* xs = smem;                                  // [PADDED]
    float* grp = xs + PADDED;                          // 4 x GRP_FLOATS
    float* tile = grp + 4 * GRP_FLOATS;                // [NMEL][TILE_LD]
    __shared__ float wmax[8], wmin[8];

    const int64_t seg = blockIdx.x;
    const int tid = threadIdx.x;
    const int g = tid >> 6;          // frame group 0..3
    const int t = tid & 63;          // thread in group

    // stage the segment, zero padded.  Segments are either rows of an (n_seg, 8000) array or -- seg_off != NULL --
    // windows of whole-track sample runs (segment s starts at sample seg_off[s] and has seg_valid[s] <= 8000 real
    // samples, the rest is the zero padding of audio_utils.py:249-252): overlapping segments are cut on the GPU
    // and every sample crosses PCIe once.
    const TIn* xin = x + (seg_off ? seg_off[seg] : seg * SEG_LEN);
    const int n_valid = seg_valid ? seg_valid[seg] : SEG_LEN;
    for (int i = tid; i < PADDED; i += 256) {
        const int s = i - PAD;
        xs[i] = (s >= 0 && s < n_valid) ? to_sample<TIn>(xin[s]) : 0.f;
    }

    // per-thread constants
    float win[16];
    float2 tw1[8], tw2[8], twp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int m0 = 2 * (t + 64 * j);
        const float2 w2 = __ldg(reinterpret_cast<const float2*>(tab + TAB_WIN + m0));
        win[2 * j] = w2.x;
        win[2 * j + 1] = w2.y;
        tw1[j] = __ldg(reinterpret_cast<const float2*>(tab + TAB_TW1) + ((t * j) & 511));
        tw2[j] = __ldg(reinterpret_cast<const float2*>(tab + TAB_TW2) + (((t >> 3) * j) & 63));
        twp[j] = __ldg(reinterpret_cast<const float2*>(tab + TAB_TWP) + (t + 64 * j));
    }

    float* P = grp + g * GRP_FLOATS;       // S (re rows 0..7, im rows 8..15, stride S_LD)  /  Z (re[512], im[512])
    float* Q = P + 2 * 8 * S_LD;           // S2 (re[64*9], im[64*9])  /  mag[512]
    float* Sre = P;
    float* Sim = P + 8 * S_LD;
    float* S2re = Q;
    float* S2im = Q + 64 * 9;
    float* Zre = P;
    float* Zim = P + 512;
    float* mag = Q;
    const int bar_id = 1 + g;
    float lmax = -INFINITY, lmin = INFINITY;
    __syncthreads();

    for (int it = 0; it < NFRAMES / 4; ++it) {
        const int frame = it * 4 + g;
        const float* xf = xs + frame * HOP;
        float2 a[8];
        // pass 1: radix-8 over n = t + 64 j
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 v = *reinterpret_cast<const float2*>(xf + 2 * (t + 64 * j));
            a[j] = make_float2(v.x * win[2 * j], v.y * win[2 * j + 1]);
        }
        fft8(a);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
            const float2 y = k1 == 0 ? a[0] : cmul(a[k1], tw1[k1]);
            Sre[k1 * S_LD + t] = y.x;
            Sim[k1 * S_LD + t] = y.y;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        // pass 2: thread (k1 = t & 7, n2a = t >> 3), radix-8 over n2b
        {
            const int k1 = t & 7, n2a = t >> 3;
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = make_float2(Sre[k1 * S_LD + n2a + 8 * j], Sim[k1 * S_LD + n2a + 8 * j]);
            fft8(a);
#pragma unroll
            for (int k2a = 0; k2a < 8; ++k2a) {
                const float2 y = k2a == 0 ? a[0] : cmul(a[k2a], tw2[k2a]);
                S2re[(k1 + 8 * k2a) * 9 + n2a] = y.x;
                S2im[(k1 + 8 * k2a) * 9 + n2a] = y.y;
            }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        // pass 3: thread t = k1 + 8 k2a, radix-8 over n2a -> Z[t + 64 k2b]
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = make_float2(S2re[t * 9 + j], S2im[t * 9 + j]);
        fft8(a);
#pragma unroll
        for (int k2b = 0; k2b < 8; ++k2b) {
            Zre[t + 64 * k2b] = a[k2b].x;
            Zim[t + 64 * k2b] = a[k2b].y;
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        // split step: X[k] = E + W^k * (-i D),  E = (Z[k] + conj Z[512-k]) / 2,  D = (Z[k] - conj Z[512-k]) / 2
        float mg[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int k = t + 64 * m;
            const int kr = (512 - k) & 511;
            const float ar = Zre[k], ai = Zim[k], br = Zre[kr], bi = -Zim[kr];
            const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
            const float dr = 0.5f * (ar - br), di = 0.5f * (ai - bi);
            const float2 o = cmul(twp[m], make_float2(di, -dr));
            const float xr = er + o.x, xi = ei + o.y;
            mg[m] = sqrtf(xr * xr + xi * xi);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // all Z / S2 reads done before mag overwrites Q
#pragma unroll
        for (int m = 0; m < 8; ++m) mag[t + 64 * m] = mg[m];
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        // sparse mel filterbank + log
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int f = t + 64 * m;
            const int st = __ldg(start + f);
            float acc = 0.f;
#pragma unroll
            for (int i = 0; i < MAXTAPS; ++i) acc += __ldg(melw + i * NMEL + f) * mag[min(st + i, 511)];
            const float y = log10f(fmaxf(acc + 0.06f, 1e-10f));
            tile[f * TILE_LD + frame] = y;
            lmax = fmaxf(lmax, y);
            lmin = fminf(lmin, y);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");     // mag reads done before next frame's S2 writes
    }

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
    }
    if ((tid & 31) == 0) { wmax[tid >> 5] = lmax; wmin[tid >> 5] = lmin; }
    __syncthreads();
    if (tid == 0) {
        float m = wmax[0], mn = wmin[0];
        for (int w = 1; w < 8; ++w) { m = fmaxf(m, wmax[w]); mn = fminf(mn, wmin[w]); }
        atomicMax(gmax + seg / group_size, f2ord(m));
        if (gmin) atomicMin(gmin + seg / group_size, f2ord(mn));      // only the segment_norm branch reads it
    }
    float* o = out + seg * (NMEL * NFRAMES);
    for (int i = tid; i < NMEL * NFRAMES / 4; i += 256) {
        const int f = (4 * i) / NFRAMES, fr = (4 * i) % NFRAMES;
        const float* src = tile + f * TILE_LD + fr;
        reinterpret_cast<float4*>(o)[i] = make_float4(src[0], src[1], src[2], src[3]);
    }
}

__global__ void logmel_fill_kernel(int32_t* p, int64_t n, int32_t v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// y = max(y - group_max, -80)   (melspectrogram.py:108-109); with gmin != NULL also the 'melspec_maxnorm' branch
// y = (y - min / 2) / |min / 2 + 1e-10| (:110-111), min = the group's minimum AFTER the subtraction and the clamp
__global__ void logmel_finish_kernel(float* __restrict__ out, int64_t n_seg, int64_t group_size,
                                     const int32_t* __restrict__ gmax, const int32_t* __restrict__ gmin) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;     // float4 index
    const int64_t per_seg = NMEL * NFRAMES / 4;
    if (i >= n_seg * per_seg) return;
    const int64_t grp = (i / per_seg) / group_size;
    const float m = ord2f(gmax[grp]);
    float4 v = reinterpret_cast<float4*>(out)[i];
    v.x = fmaxf(v.x - m, -80.f);
    v.y = fmaxf(v.y - m, -80.f);
    v.z = fmaxf(v.z - m, -80.f);
    v.w = fmaxf(v.w - m, -80.f);
    if (gmin) {
        const float half = fmaxf(ord2f(gmin[grp]) - m, -80.f) / 2.f;
        const float den = fabsf(half + 1e-10f);
        v.x = (v.x - half) / den;
        v.y = (v.y - half) / den;
        v.z = (v.z - half) / den;
        v.w = (v.w - half) / den;
    }
    reinterpret_cast<float4*>(out)[i] = v;
}

// ---- host: librosa-0.8.1-style Slaney mel filterbank (htk=False, norm='slaney'), in double
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static int logmel_init(nafp_ctx* ctx) {
    if (ctx->logmel) return NAFP_OK;
    const double fs = 8000.0, fmin = 300.0, fmax = 4000.0;
    const int n_freq = NFFT / 2 + 1;
    std::vector<double> mel_f(NMEL + 2);
    const double m_lo = hz_to_mel(fmin), m_hi = hz_to_mel(fmax);
    for (int i = 0; i < NMEL + 2; ++i) mel_f[i] = mel_to_hz(m_lo + (m_hi - m_lo) * i / (NMEL + 1));
    std::vector<float> w(MAXTAPS * NMEL, 0.f);
    std::vector<int> st(NMEL, 0);
    for (int f = 0; f < NMEL; ++f) {
        const double enorm = 2.0 / (mel_f[f + 2] - mel_f[f]);
        int first = -1, count = 0;
        for (int b = 0; b < n_freq; ++b) {
            const double fr = (fs / 2.0) * b / (n_freq - 1);
            const double lower = (fr - mel_f[f]) / (mel_f[f + 1] - mel_f[f]);
            const double upper = (mel_f[f + 2] - fr) / (mel_f[f + 2] - mel_f[f + 1]);
            const double v = std::fmax(0.0, std::fmin(lower, upper)) * enorm;
            if (static_cast<float>(v) > 0.f) {
                if (first < 0) first = b;
                NAFP_REQUIRE(b - first < MAXTAPS && b <= 511, NAFP_ERR_UNSUPPORTED,
                             "logmel: mel band %d has a tap outside the %d-tap / 511-bin layout", f, MAXTAPS);
                w[(b - first) * NMEL + f] = static_cast<float>(v);
                ++count;
            }
        }
        NAFP_REQUIRE(first >= 0 && count > 0, NAFP_ERR_UNSUPPORTED, "logmel: empty mel band %d", f);
        st[f] = first;
    }
    LogmelState* s = new LogmelState();
    {
        std::vector<float> tab(TAB_FLOATS);
        const double pi = 3.14159265358979323846;
        for (int i = 0; i < 1024; ++i) tab[TAB_WIN + i] = static_cast<float>(0.5 - 0.5 * std::cos(2.0 * pi * i / 1024.0));
        auto fill = [&](int off, int den, int count) {          // e^{-2 pi i k / den}, k < count
            for (int k = 0; k < count; ++k) {
                tab[off + 2 * k] = static_cast<float>(std::cos(2.0 * pi * k / den));
                tab[off + 2 * k + 1] = static_cast<float>(-std::sin(2.0 * pi * k / den));
            }
        };
        fill(TAB_TW1, 512, 512);
        fill(TAB_TW2, 64, 64);
        fill(TAB_TWP, 1024, 512);
        NAFP_CUDA(cudaMalloc(&s->tab, tab.size() * sizeof(float)));
        NAFP_CUDA(cudaMemcpy(s->tab, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    NAFP_CUDA(cudaMalloc(&s->melw, w.size() * sizeof(float)));
    NAFP_CUDA(cudaMalloc(&s->start, st.size() * sizeof(int)));
    NAFP_CUDA(cudaMemcpy(s->melw, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice));
    NAFP_CUDA(cudaMemcpy(s->start, st.data(), st.size() * sizeof(int), cudaMemcpyHostToDevice));
    NAFP_CUDA(cudaFuncSetAttribute(logmel_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOGMEL_SMEM));
    NAFP_CUDA(cudaFuncSetAttribute(logmel_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, LOGMEL_SMEM));
    ctx->logmel = s;
    return NAFP_OK;
}

void logmel_destroy(nafp_ctx* ctx) {
    if (!ctx->logmel) return;
    cudaFree(ctx->logmel->melw);
    cudaFree(ctx->logmel->start);
    cudaFree(ctx->logmel->tab);
    if (ctx->logmel->gmax) cudaFree(ctx->logmel->gmax);
    delete ctx->logmel;
    ctx->logmel = nullptr;
}

// raw (un-normalised) log-mel + per-group maxima; `finish` applies the max subtraction and clamp.
// Returns the device pointer of the group maxima through gmax_out (valid until the next call).
int logmel_run(nafp_ctx* ctx, const void* x_dev, bool pcm16, int64_t n_seg, int64_t group_size, float* mel_dev,
               bool finish, const int32_t** gmax_out, const int64_t* seg_off, const int32_t* seg_valid) {
    NAFP_REQUIRE(ctx && (n_seg == 0 || (x_dev && mel_dev)) && n_seg >= 0 && group_size >= 1, NAFP_ERR_INVALID,
                 "logmel: bad arguments (n_seg=%lld group_size=%lld)", (long long)n_seg, (long long)group_size);
    NAFP_CUDA(cudaSetDevice(ctx->device));
    NAFP_TRY(logmel_init(ctx));
    if (n_seg == 0) return NAFP_OK;
    LogmelState* s = ctx->logmel;
    const int64_t groups = (n_seg + group_size - 1) / group_size;
    if (s->gmax_cap < groups) {
        if (s->gmax) NAFP_CUDA(cudaFree(s->gmax));
        s->gmax = nullptr;
        s->gmax_cap = 0;
        NAFP_CUDA(cudaMalloc(&s->gmax, 2 * static_cast<size_t>(groups) * sizeof(int32_t)));
        s->gmax_cap = groups;
    }
    int32_t* gmin = s->segment_norm ? s->gmax + s->gmax_cap : nullptr;
    NAFP_REQUIRE(finish || !s->segment_norm, NAFP_ERR_STATE, "logmel: melspec_maxnorm needs the finishing pass");
    logmel_fill_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, ctx->stream>>>(
        s->gmax, groups, static_cast<int32_t>(0x807FFFFFu));                 // f2ord(-inf)
    if (gmin) {
        logmel_fill_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, ctx->stream>>>(
            gmin, groups, static_cast<int32_t>(0x7F800000u));                // f2ord(+inf)
        ctx->launches++;
    }
    if (pcm16)
        logmel_kernel<int16_t><<<static_cast<unsigned>(n_seg), 256, LOGMEL_SMEM, ctx->stream>>>(
            static_cast<const int16_t*>(x_dev), n_seg, group_size, s->melw, s->start, s->tab, mel_dev, s->gmax, gmin, seg_off, seg_valid);
    else
        logmel_kernel<float><<<static_cast<unsigned>(n_seg), 256, LOGMEL_SMEM, ctx->stream>>>(
            static_cast<const float*>(x_dev), n_seg, group_size, s->melw, s->start, s->tab, mel_dev, s->gmax, gmin, seg_off, seg_valid);
    ctx->launches += 2;
    if (finish) {
        const int64_t n4 = n_seg * (NMEL * NFRAMES / 4);
        logmel_finish_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, ctx->stream>>>(mel_dev, n_seg, group_size,
                                                                                             s->gmax, gmin);
        ctx->launches++;
    }
    NAFP_CUDA(cudaGetLastError());
    if (gmax_out) *gmax_out = s->gmax;
    return NAFP_OK;
}

bool logmel_segment_norm(nafp_ctx* ctx) { return ctx->logmel && ctx->logmel->segment_norm; }

}  // namespace nafp

/* MODEL.FEAT: 0 = 'melspec' (default), 1 = 'melspec_maxnorm' (Melspec_layer(segment_norm=True), melspectrogram.py:110-111) */
extern "C" int nafp_logmel_set_segment_norm(nafp_ctx* ctx, int32_t enable) {
    NAFP_REQUIRE(ctx, NAFP_ERR_INVALID, "nafp_logmel_set_segment_norm: ctx is NULL");
    NAFP_CUDA(cudaSetDevice(ctx->device));
    NAFP_TRY(nafp::logmel_init(ctx));
    NAFP_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->logmel->segment_norm = enable != 0;
    return NAFP_OK;
}

extern "C" int nafp_logmel_forward(nafp_ctx* ctx, const float* x_dev, int64_t n_seg, int64_t group_size,
                                   float* mel_dev) {
    NAFP_RANGE("nafp_logmel_forward");
    NAFP_REQUIRE(ctx, NAFP_ERR_INVALID, "nafp_logmel_forward: ctx is NULL");
    return nafp::logmel_run(ctx, x_dev, false, n_seg, group_size, mel_dev, true, nullptr, nullptr, nullptr);
}

/* The log-mel kernel alone, as the fused extractor runs it: raw log10(mel + 0.06) + the per-group maxima (ordered-int
 * encoding of the float maximum; NULL to skip the copy); the "- max, clamp" of melspectrogram.py:108-109 is then applied by
 * the encoder's first layer.  bench.py times this entry for the log-mel roofline (64,768 algorithmic bytes per segment). */
extern "C" int nafp_logmel_forward_raw(nafp_ctx* ctx, const float* x_dev, int64_t n_seg, int64_t group_size, float* mel_dev,
                                       int32_t* group_max_dev) {
    NAFP_RANGE("nafp_logmel_forward_raw");
    NAFP_REQUIRE(ctx, NAFP_ERR_INVALID, "nafp_logmel_forward_raw: ctx is NULL");
    NAFP_REQUIRE(!nafp::logmel_segment_norm(ctx), NAFP_ERR_STATE, "nafp_logmel_forward_raw: melspec_maxnorm needs the finishing pass");
    const int32_t* gmax = nullptr;
    NAFP_TRY(nafp::logmel_run(ctx, x_dev, false, n_seg, group_size, mel_dev, false, &gmax, nullptr, nullptr));
    if (group_max_dev && n_seg > 0)
        NAFP_CUDA(cudaMemcpyAsync(group_max_dev, gmax, static_cast<size_t>((n_seg + group_size - 1) / group_size) * sizeof(int32_t),
                                  cudaMemcpyDeviceToDevice, ctx->stream));
    return NAFP_OK;
}
