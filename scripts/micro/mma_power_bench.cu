// Microbenchmark: sustained tcgen05.mma rate (M=128, N=256, K=16, bf16 -> fp32) with both operands from shared
// memory ("SS", what the scan does) versus the A operand resident in tensor memory ("TS"), all 148 SMs busy for
// ~0.3 s so that the board reaches its power equilibrium.  Operand contents are irrelevant here.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../neural-audio-fp_b200/csrc/ptx.cuh"
using namespace nafp;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}

template <bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    // random bf16 values in [-2, 2): realistic operand toggling (constant operands draw far less power)
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x807f807fu) | 0x3f803f80u;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    {   // random A operand in tensor memory: columns 256..287 of every lane
        uint32_t r[8];
        for (int c0 = 0; c0 < 32; c0 += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t h = (threadIdx.x * 64 + c0 + j + 7) * 2654435761u;
                h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                r[j] = (h & 0x807f807fu) | 0x3f803f80u;
            }
            const uint32_t taddr = tb + (static_cast<uint32_t>(warp * 32) << 16) + 256 + c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                         "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before(); __syncthreads(); tc_fence_after();
    }
    if (warp == 0) {
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_f16(1u, 128u, 256u);
        const uint64_t adesc = umma_desc_sw128(smem_u32(smem));              // 128 rows x 64 cols (16 KB)
        const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + 16384));      // 256 rows x 64 cols (32 KB)
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t d = tb + (j & 1) * 256 * 0;               // one accumulator
                    if (TS) mma_ts(d, tb + 256 + (j & 3) * 8, bdesc + 2 * (j & 3), idesc, (j & 3) ? 1u : 0u);
                    else    tc_mma_f16(d, adesc + 2 * (j & 3), bdesc + 2 * (j & 3), idesc, (j & 3) ? 1u : 0u);
                }
                tc_commit(&bar);
            }
            __syncwarp();
            mbar_wait_parked(&bar, it & 1);
        }
        if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <bool TS> void run(const char* name, long long* cyc) {
    cudaFuncSetAttribute(k<TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 120000;               // x 32 MMAs x 128 cycles ~ 0.3 s
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<TS><<<148, 128, 50 * 1024>>>(iters, cyc);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 148.0 * iters * 32.0 * 2.0 * 128 * 256 * 16;
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%s rep %d: %.1f ms, %.0f TFLOP/s, %.1f cycles per MMA, effective clock %.2f GHz\n", name, rep, ms, flop / ms / 1e9,
               double(h[0]) / iters / 32.0, double(h[0]) / (ms * 1e6));
    }
}

int main() {
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    run<false>("SS (A and B from shared memory)", cyc);
    run<true>("TS (A from tensor memory)     ", cyc);
    run<false>("SS again                      ", cyc);
    return 0;
}
