// Micro-benchmark (diagnostics, not product): cycles per tcgen05.mma kind::f16 instruction as a function of N, of where
// the A operand lives (tensor memory vs shared memory), of M, and of how many independent accumulators the
// stream of MMAs rotates over.  Answers "what does one M128 N16 K16 MMA really cost in the recurrent kernel".
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I medaka_b200/csrc tools/umma_bench.cu -o build/umma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace mdk;

constexpr int KSTEPS = 8;
constexpr int SMEM_BYTES = 96 * 1024;

// MODE 0: A in TMEM (.ts), MODE 1: A in smem (.ss)
template <int MODE, int M, int N, int NACC, int ISSUERS, int PAD = 0, int POLLERS = 0>
__global__ void __launch_bounds__(32 * (4 + ISSUERS + POLLERS), 1) bench_kernel(long long *out, int reps) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + SMEM_BYTES - 64);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + SMEM_BYTES - 16);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (SMEM_BYTES - 64) / 16; i += blockDim.x) reinterpret_cast<int4 *>(smem)[i] = make_int4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(bar, ISSUERS);
        mbar_init(bar + 1, ISSUERS);
        fence_mbar_init();
    }
    if (warp == 4) {
        tmem_alloc(slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp < 4) {
        // zero the TMEM A region (columns 0..255) so the MMAs run on finite data
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 256; c += 8) tmem_st_x8(((uint32_t)(warp * 32) << 16) + c, z);
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp >= 4 + ISSUERS) {
        // pollers: sit in mbarrier.try_wait like the recurrent kernel's gate warps do during the MMA phase
        mbar_wait(bar + 1, 0);
    } else if (warp >= 4) {
        const int g = warp - 4;
        constexpr uint32_t idesc = make_idesc_f16(M, N);
        // B: [kg][N rows][16 B]; A (smem): [kg][M rows][16 B] at +32 KiB
        const uint64_t b0 = make_smem_desc(smem_u32(smem), N * 16 + PAD, 128);
        const uint64_t a0 = make_smem_desc(smem_u32(smem + 32768), M * 16, 128);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int i = 0; i < 24; ++i) {
                    const int ks = i % KSTEPS;
                    const uint32_t d = 256u + (uint32_t)(((g * NACC + (i % NACC)) * N) % 256);
                    const uint64_t bd = b0 + (uint64_t)((ks * 2 * (N * 16 + PAD)) >> 4);
                    if (MODE == 0)
                        umma_f16_ts(d, (uint32_t)(((g * 3 + i / KSTEPS) * 8 + ks) * 8 % 256), bd, idesc, 1u);
                    else
                        umma_f16(d, a0 + (uint64_t)((ks * 2 * M * 16) >> 4), bd, idesc, 1u);
                }
            }
            t1 = clock64();
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t2 = clock64();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bar + 1);
        if (g == 0 && (threadIdx.x & 31) == 0 && blockIdx.x == 0) {
            out[0] = t1 - t0;   // only meaningful if lane 0 was the elected lane (it is, in practice)
            out[1] = t2 - t0;
        }
        if (elect_one() && g == 0 && blockIdx.x == 0) {
            out[2] = t1 - t0;
            out[3] = t2 - t0;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after_sync();
        tmem_dealloc(0, 512);
    }
}

template <int MODE, int M, int N, int NACC, int ISSUERS, int PAD = 0, int POLLERS = 0>
void run(const char *name, int grid) {
    long long *d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    auto k = bench_kernel<MODE, M, N, NACC, ISSUERS, PAD, POLLERS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    const int reps = 8;
    long long h[4];
    for (int it = 0; it < 2; ++it) {
        k<<<grid, 32 * (4 + ISSUERS + POLLERS), SMEM_BYTES>>>(d, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
            exit(1);
        }
    }
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    const int n = 24 * reps * ISSUERS;
    printf("%-44s grid %3d  MMAs %4d  issue %6lld cyc  done %6lld cyc  -> %6.2f cyc/MMA (issue %5.2f)\n", name, grid, n, h[2],
           h[3], (double)h[3] / n, (double)h[2] / (24 * reps));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<0, 128, 16, 1, 3>("TS N16 3 issuers", grid);
        run<0, 128, 16, 1, 2>("TS N16 2 issuers", grid);
        run<0, 128, 16, 1, 3, 16>("TS N16 3 issuers, LBO 272", grid);
        run<0, 128, 16, 1, 3, 0, 12>("TS N16 3 issuers, 12 polling warps", grid);
        run<0, 128, 16, 1, 3, 16, 12>("TS N16 3 issuers, LBO 272, 12 pollers", grid);
        run<0, 128, 16, 1, 2, 16, 12>("TS N16 2 issuers, LBO 272, 12 pollers", grid);
        run<0, 128, 32, 1, 3, 16>("TS N32 3 issuers, LBO 528", grid);
        run<0, 128, 8, 1, 3, 16>("TS N8  3 issuers, LBO 144", grid);
    }
    return 0;
}
