// Microbenchmark: HBM read bandwidth of a TMA stream as a function of the BYTES IN FLIGHT per SM.
//
// The flat scan streams its database through a 4-slot ring of 32 KB K-block slots and tops out at 4.7-5.1 TB/s
// whatever the SM clock: the hypothesis is that one 64 KB tile in flight per SM against
// ~2 us of loaded memory latency is what bounds it.  This kernel has the scan's access pattern and nothing else:
// one producer thread per CTA issues the same 128-row x 64-column bf16 boxes (16 KB, SWIZZLE_128B) of a row-major
// (rows, 128) bf16 matrix into a ring of S slots, one consumer thread waits for each slot and hands it straight back.
// S x 16 KB is then the number of bytes a CTA can have in flight.  Output: GB/s for S = 2 .. 13.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_inflight_bench tma_inflight_bench.cu && ./tma_inflight_bench
// NOT RUN YET (written at the end of round 1 with the GPU budget spent) -- first thing to run in round 2.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <cstdint>
#include <cstdio>

#include "../../neural-audio-fp_b200/csrc/ptx.cuh"
using namespace nafp;

constexpr int SLOT_BYTES = 128 * 128;      // 128 rows x 64 bf16
constexpr int MAX_SLOTS = 13;

__global__ void __launch_bounds__(64, 1)
stream_kernel(const __grid_constant__ CUtensorMap tmap, int n_tiles, int slots, unsigned long long* sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[MAX_SLOTS], empty[MAX_SLOTS];
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < slots; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item = (128-row tile, K block): the two boxes of a tile are consecutive items, as in the scan
    const int n_items = 2 * ((n_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x));
    if (warp == 0 && lane == 0) {
        for (int it = 0; it < n_items; ++it) {
            const int s = it % slots;
            const uint32_t ph = (it / slots) & 1;
            mbar_wait_parked(&empty[s], ph ^ 1);
            const int tile = static_cast<int>(blockIdx.x) + (it >> 1) * static_cast<int>(gridDim.x);
            mbar_arrive_expect_tx(&full[s], SLOT_BYTES);
            tma_load_2d(smem + s * SLOT_BYTES, &tmap, &full[s], (it & 1) * 64, tile * 128);
        }
    } else if (warp == 1 && lane == 0) {
        unsigned long long acc = 0;
        for (int it = 0; it < n_items; ++it) {
            const int s = it % slots;
            const uint32_t ph = (it / slots) & 1;
            mbar_wait_parked(&full[s], ph);
            acc += *reinterpret_cast<const volatile unsigned long long*>(smem + s * SLOT_BYTES);     // touch the slot
            mbar_arrive(&empty[s]);
        }
        sink[blockIdx.x] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const long long rows = argc > 1 ? atoll(argv[1]) : 16000000ll;        // 4 GB of bf16 rows by default (>> L2)
    const int n_tiles = static_cast<int>(rows / 128);
    void* db = nullptr;
    if (cudaMalloc(&db, static_cast<size_t>(n_tiles) * 128 * 256) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
    cudaMemset(db, 0x3c, static_cast<size_t>(n_tiles) * 128 * 256);
    unsigned long long* sink;
    cudaMalloc(&sink, 148 * 8);

    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
        printf("cuTensorMapEncodeTiled entry point unavailable\n");
        return 1;
    }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {128, static_cast<cuuint64_t>(n_tiles) * 128};
    const cuuint64_t gstr[1] = {256};
    const cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    const CUresult r = reinterpret_cast<EncodeFn>(fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, gdim, gstr, box, es,
                                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r)); return 1; }

    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int smem_bytes = MAX_SLOTS * SLOT_BYTES + 1024;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = static_cast<double>(n_tiles) * 128 * 256;
    for (int slots = 2; slots <= MAX_SLOTS; ++slots) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            stream_kernel<<<sms, 64, smem_bytes>>>(tmap, n_tiles, slots, sink);
            cudaEventRecord(e1);
            const cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("slots %d: %s\n", slots, cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        printf("slots %2d (%3d KB in flight per SM): %.3f ms, %.0f GB/s\n", slots, slots * SLOT_BYTES / 1024, best, bytes / best / 1e6);
    }
    return 0;
}
