// Synthetic inputs generated on the device (SURVEY §8 d): datasets are unavailable offline and the
// full-scale shapes (56 M fingerprints = 28.7 GB, 56 M one-second segments = 1.8 TB) cannot be staged
// from the host, so every shard regenerates its own slice from a counter-based generator keyed by
// (seed, global row / segment id).  Measurement and test infrastructure, not part of the reference path.
#include "common.h"

namespace nafp {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t hash4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return mix32(a ^ mix32(b + 0x9e3779b9u ^ mix32(c + 0x85ebca6bu ^ mix32(d + 0xc2b2ae35u))));
}
__device__ __forceinline__ float u01(uint32_t h) { return (static_cast<float>(h >> 8) + 0.5f) * (1.0f / 16777216.0f); }
// two standard normals from two hashes
__device__ __forceinline__ float2 normal2(uint32_t h0, uint32_t h1) {
    const float r = sqrtf(-2.0f * logf(u01(h0)));
    float s, c;
    sincospif(2.0f * u01(h1), &s, &c);
    return make_float2(r * c, r * s);
}

// unit-norm 128-d rows, AR(1) (rho) along the rows of one `track_len`-row track; one warp per track
__global__ void synth_fp_rows_kernel(uint32_t seed, int64_t row0, int64_t n_rows, int track_len, float rho,
                                     float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t first_track = row0 / track_len;
    const int64_t last_track = (row0 + n_rows - 1) / track_len;
    const int64_t track = first_track + warp;
    if (track > last_track) return;
    const float c = sqrtf(1.0f - rho * rho);
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < track_len; ++j) {
        const uint32_t t_lo = static_cast<uint32_t>(track), t_hi = static_cast<uint32_t>(track >> 32);
        const float2 a = normal2(hash4(seed, t_lo, t_hi * 64 + j, 8 * lane + 0), hash4(seed, t_lo, t_hi * 64 + j, 8 * lane + 1));
        const float2 b = normal2(hash4(seed, t_lo, t_hi * 64 + j, 8 * lane + 2), hash4(seed, t_lo, t_hi * 64 + j, 8 * lane + 3));
        const float e[4] = {a.x, a.y, b.x, b.y};
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            r[k] = j == 0 ? e[k] : rho * r[k] + c * e[k];
            ss += r[k] * r[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = rsqrtf(ss);
        const int64_t row = track * track_len + j;
        if (row >= row0 && row < row0 + n_rows)
            reinterpret_cast<float4*>(out + (row - row0) * 128)[lane] = make_float4(r[0] * inv, r[1] * inv, r[2] * inv, r[3] * inv);
    }
}

// one-second 8 kHz segments: a few amplitude-modulated tones in 300..3800 Hz + white noise, peak ~0.5
__global__ void synth_audio_kernel(uint32_t seed, int64_t seg0, int64_t n_seg, float* __restrict__ out) {
    const int64_t seg = blockIdx.x;
    if (seg >= n_seg) return;
    const uint32_t id_lo = static_cast<uint32_t>(seg0 + seg), id_hi = static_cast<uint32_t>((seg0 + seg) >> 32);
    __shared__ float f[8], ph[8], amp[8], amf[8];
    if (threadIdx.x < 8) {
        const int k = threadIdx.x;
        f[k] = 300.f + 3500.f * u01(hash4(seed, id_lo, id_hi, 1000 + k));
        ph[k] = 2.f * u01(hash4(seed, id_lo, id_hi, 2000 + k));
        amp[k] = 0.3f + 0.7f * u01(hash4(seed, id_lo, id_hi, 3000 + k));
        amf[k] = 0.2f + 3.8f * u01(hash4(seed, id_lo, id_hi, 4000 + k));
    }
    __syncthreads();
    for (int n = threadIdx.x; n < 8000; n += blockDim.x) {
        const float t = static_cast<float>(n) * (1.0f / 8000.0f);
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            v += amp[k] * (0.6f + 0.4f * sinpif(2.f * amf[k] * t)) * sinpif(2.f * f[k] * t + ph[k]);
        const float2 g = normal2(hash4(seed, id_lo, id_hi * 8192 + n, 7), hash4(seed, id_lo, id_hi * 8192 + n, 11));
        out[seg * 8000 + n] = 0.1f * v + 0.02f * g.x;
    }
}

}  // namespace nafp

using namespace nafp;

extern "C" {

int nafp_synth_fp_rows(nafp_ctx* ctx, int64_t seed, int64_t row0, int64_t n_rows, int32_t track_len, float rho,
                       float* out_dev) {
    NAFP_RANGE("nafp_synth_fp_rows");
    NAFP_REQUIRE(ctx && out_dev && n_rows >= 0 && row0 >= 0 && track_len >= 1 && track_len <= 64, NAFP_ERR_INVALID,
                 "nafp_synth_fp_rows: bad arguments (track_len <= 64)");
    if (n_rows == 0) return NAFP_OK;
    NAFP_CUDA(cudaSetDevice(ctx->device));
    const int64_t tracks = (row0 + n_rows - 1) / track_len - row0 / track_len + 1;
    synth_fp_rows_kernel<<<static_cast<unsigned>((tracks * 32 + 255) / 256), 256, 0, ctx->stream>>>(
        static_cast<uint32_t>(seed), row0, n_rows, track_len, rho, out_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

int nafp_synth_audio(nafp_ctx* ctx, int64_t seed, int64_t seg0, int64_t n_seg, float* out_dev) {
    NAFP_RANGE("nafp_synth_audio");
    NAFP_REQUIRE(ctx && out_dev && n_seg >= 0 && seg0 >= 0, NAFP_ERR_INVALID, "nafp_synth_audio: bad arguments");
    if (n_seg == 0) return NAFP_OK;
    NAFP_CUDA(cudaSetDevice(ctx->device));
    synth_audio_kernel<<<static_cast<unsigned>(n_seg), 256, 0, ctx->stream>>>(static_cast<uint32_t>(seed), seg0, n_seg, out_dev);
    ctx->launches++;
    NAFP_CUDA(cudaGetLastError());
    return NAFP_OK;
}

}  // extern "C"
