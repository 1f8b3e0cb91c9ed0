// Micro-test: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st), B in shared memory (SWIZZLE_128B,
// K-major), against a CPU product.  Establishes the TMEM layout of a 16-bit A operand before the flat scan keeps its
// query tile there: lane = row m, 32-bit column c holds elements k = 2c (low half) and k = 2c + 1 (high half).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_a_test tmem_a_test.cu && ./tmem_a_test
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../neural-audio-fp_b200/csrc/ptx.cuh"
using namespace nafp;

constexpr int M = 128, N = 128, K = 64;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __launch_bounds__(128, 1)
test_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tbase;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tbase, 256); tmem_relinquish(); }
    // B tile (N rows x 64 bf16 = 128 B per row) into shared memory with the 128-byte swizzle by hand
    for (int i = threadIdx.x; i < N * 8; i += blockDim.x) {
        const int n = i >> 3, chunk = i & 7;                    // 16-byte chunk of row n
        const uint4 v = reinterpret_cast<const uint4*>(B + n * K)[chunk];
        *reinterpret_cast<uint4*>(smem + (n >> 3) * 1024 + (n & 7) * 128 + ((chunk ^ (n & 7)) << 4)) = v;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tbase;
    // A: thread = row m; 64 bf16 = 32 words into TMEM columns [128, 160) of its lane
    {
        const int m = warp * 32 + lane;
        uint32_t v[16];
        for (int h = 0; h < 2; ++h) {
            for (int j = 0; j < 16; ++j) v[j] = reinterpret_cast<const uint32_t*>(A + m * K)[h * 16 + j];
            tmem_st_32x16(tb + (static_cast<uint32_t>(warp * 32) << 16) + 128 + h * 16, v);
        }
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc_f16(1u, M, N);
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem));
            for (int j = 0; j < K / 16; ++j) mma_ts(tb, tb + 128 + j * 8, bdesc + 2 * j, idesc, j ? 1u : 0u);
            tc_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    {
        const int m = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tb + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
            tc_wait_ld();
            for (int j = 0; j < 32; ++j) D[m * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 256);
}

int main() {
    std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
    std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * K; ++i) { hB[i] = __float2bfloat16(rnd()); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += static_cast<double>(fA[m * K + k]) * fB[n * K + k];
            ref[m * N + n] = static_cast<float>(acc);
        }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    test_kernel<<<1, 128, 32 * 1024>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int i = 0; i < M * N; ++i) worst = fmax(worst, fabs(out[i] - ref[i]));
    printf("tcgen05.mma A-from-TMEM (M %d, N %d, K %d): max |err| vs CPU = %.3e  -> %s\n", M, N, K, worst, worst < 1e-3 ? "LAYOUT OK" : "MISMATCH");
    return worst < 1e-3 ? 0 : 2;
}
