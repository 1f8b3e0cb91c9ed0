// Microbenchmark: tcgen05.ld (LDTM) latency / throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu && ./ldtm_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../neural-audio-fp_b200/csrc/ptx.cuh"
using namespace nafp;

__global__ void k(int nwarps, int reps, int mode, long long* out, uint32_t* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) { tmem_alloc(&tbase, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < nwarps) {
        if (mode == 0) {            // ld + wait each time (latency-bound chain)
            for (int r = 0; r < reps; ++r) {
                uint32_t v[32]; tmem_ld_32x32(base + ((r * 32) & 255) + (warp >> 2) * 0, v); tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += v[j];
            }
        } else if (mode == 1) {     // two loads in flight
            for (int r = 0; r < reps; r += 2) {
                uint32_t v[32], w[32];
                tmem_ld_32x32(base + ((r * 32) & 255), v); tmem_ld_32x32(base + (((r + 1) * 32) & 255), w); tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += v[j] ^ w[j];
            }
        } else {                    // x16 loads
            for (int r = 0; r < reps; ++r) {
                uint32_t v[16]; tmem_ld_32x16(base + ((r * 16) & 255), v); tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc += v[j];
            }
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (lane == 0) out[blockIdx.x * 32 + warp] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

int main() {
    long long* out; uint32_t* sink;
    cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 148 * 1024 * 4);
    long long h[32];
    for (int mode = 0; mode < 3; ++mode)
        for (int nw : {1, 4, 8, 16}) {
            const int reps = 256;
            k<<<148, 512, 0>>>(nw, reps, mode, out, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            double cyc = 0; for (int w = 0; w < nw; ++w) cyc = h[w] > cyc ? h[w] : cyc;
            const double bytes = (double)nw * reps * (mode == 2 ? 2048 : 4096);
            printf("mode %d warps %2d: %8.0f cycles for %d loads/warp -> %.1f cyc/load/warp, %.1f B/cycle/SM\n", mode, nw, cyc, reps,
                   cyc / reps, bytes / cyc);
        }
    return 0;
}
