// Tensor-core (tcgen05) implementation of the GRU gate matmuls for sm_100a  (MDK_PREC_TC).
//
// Reference arithmetic: torch.nn.GRU as used by medaka/architectures/gru.py:46-52,66; parity target is
// the fp32 CPU path (medaka/prediction.py:146-148).  To stay inside 1e-3 (scale-aware) of fp32 through a
// 10 000-step recurrence, every operand is carried as an fp16 pair (hi + lo, ~22 significant bits) and each
// product is three kind::f16 MMAs, hi*hi + hi*lo + lo*hi, accumulated in fp32 in TMEM.
//
// Both kernels use the TRANSPOSED formulation  G^T[gate rows, N] = W[gate rows, K] . X^T[K, N]:
//   A operand = weight block (M = 128 gate rows, K-major = torch's native [out][in] layout), resident in smem
//   B operand = activations (N windows or positions, K-major = row-major [n][k])
//   D (TMEM)  = lane j <-> hidden unit j, column n <-> window/position n
// so the thread that owns TMEM lane j reads r_j, z_j, n_j of one window from three column groups with no
// cross-thread exchange, and its global stores are 128-byte coalesced over j.
//
// Shared-memory operand layout (K-major, SWIZZLE_NONE): [k-group = k/8][row][8 halfs]; a core matrix is 8 rows
// x 16 B = 128 contiguous bytes, LBO = rows*16 B (next k-group), SBO = 128 B (next 8 rows).
#include "common.cuh"
#include "ptx.cuh"

namespace mdk {

__device__ __forceinline__ float sigmoid_fast(float x) {
    // 1/(1+e^-x) with ex2.approx + rcp.approx: ~2 ulp, abs error < 2e-7
    return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float tanh_fast(float x) {
    // 1 - 2/(1+e^{2x}); abs error ~1e-7 (no cancellation blow-up in absolute terms); saturates cleanly
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, 1.0f + e);
}

// =====================================================================================================
// Recurrent kernel.  One CTA = NT tiles of 16 windows of one direction, for the whole sequence.
//   warps 0-7 : gate warps (TMEM -> registers -> gate math -> next h into smem + global output)
//   warp  8   : MMA issuer (lane 0) and TMEM owner
// NT == 2: warpgroup g owns tile g; the MMA of one tile overlaps the gate math of the other (ping-pong).
// NT == 1: both warpgroups share tile 0 (8 windows each): used when there are too few windows to fill the GPU.
// =====================================================================================================
constexpr int RT_N = 16;                                 // windows per tile (UMMA N)
constexpr int RT_W_BYTES = 2 * 3 * H * H * 2;            // W_hh hi+lo, 3 gate blocks: 196 608 B
constexpr int RT_HPLANE = RT_N * H * 2;                  // one h plane (hi or lo) of a tile: 4096 B
constexpr int RT_THREADS = 288;
constexpr uint32_t RT_TMEM_COLS = 128;

template <int NT>
struct RecSmem {
    static constexpr int w_off = 0;
    static constexpr int h_off = RT_W_BYTES;                       // [NT][2 planes][4096]
    static constexpr int bar_off = h_off + NT * 2 * RT_HPLANE;     // acc_ready[NT], h_ready[NT]
    static constexpr int tmem_off = bar_off + 2 * NT * 8;
    static constexpr int total = tmem_off + 16;
};

template <int NT, bool OUT_TILES>
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_kernel(const float *__restrict__ gi, const __half *__restrict__ w_hh_tc, const float *__restrict__ b_hn,
              void *__restrict__ h_out, int64_t B, int64_t T) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = RecSmem<NT>;
    uint64_t *acc_ready = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint64_t *h_ready = acc_ready + NT;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::tmem_off);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int dir = blockIdx.y;
    const int64_t win0 = (int64_t)blockIdx.x * (RT_N * NT);

    // ---- prologue: weights -> smem (generic proxy), zero h tiles, barriers, TMEM ----
    {
        const int4 *src = reinterpret_cast<const int4 *>(reinterpret_cast<const uint8_t *>(w_hh_tc) +
                                                         (size_t)dir * RT_W_BYTES);
        int4 *dst = reinterpret_cast<int4 *>(smem + L::w_off);
        for (int i = tid; i < RT_W_BYTES / 16; i += RT_THREADS) dst[i] = src[i];
        int4 *hz = reinterpret_cast<int4 *>(smem + L::h_off);
        for (int i = tid; i < NT * 2 * RT_HPLANE / 16; i += RT_THREADS) hz[i] = make_int4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int i = 0; i < NT; ++i) {
            mbar_init(&acc_ready[i], 1);
            mbar_init(&h_ready[i], NT == 2 ? 128 : 256);
        }
        fence_mbar_init();
    }
    if (warp == 8) {
        tmem_alloc(tmem_slot, RT_TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc_f16(128, RT_N);
        const uint32_t w_addr = smem_u32(smem + L::w_off);
        const uint32_t h_addr = smem_u32(smem + L::h_off);
        for (int64_t step = 0; step < T; ++step) {
            const uint32_t par = (uint32_t)(step & 1);
#pragma unroll
            for (int tile = 0; tile < NT; ++tile) {
                mbar_wait(&h_ready[tile], par);
                tc_fence_after_sync();
                if (lane == 0) {
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const uint32_t d = tmem_base + (uint32_t)(tile * 48 + g * 16);
                        uint32_t acc = 0;
#pragma unroll
                        for (int prod = 0; prod < 3; ++prod) {
                            const int pa = (prod == 2) ? 1 : 0;   // W part: hi, hi, lo
                            const int pb = (prod == 1) ? 1 : 0;   // h part: hi, lo, hi
                            const uint32_t a0 = w_addr + (uint32_t)((pa * 3 + g) * (H * H * 2));
                            const uint32_t b0 = h_addr + (uint32_t)((tile * 2 + pb) * RT_HPLANE);
#pragma unroll
                            for (int ks = 0; ks < H / 16; ++ks) {
                                const uint64_t ad = make_smem_desc(a0 + ks * 2 * (H * 16), H * 16, 128);
                                const uint64_t bd = make_smem_desc(b0 + ks * 2 * (RT_N * 16), RT_N * 16, 128);
                                umma_f16(d, ad, bd, idesc, acc);
                                acc = 1;
                            }
                        }
                    }
                    umma_commit(&acc_ready[tile]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================= gate warps =================
        const int wg = warp >> 2;
        const int tile = (NT == 2) ? wg : 0;
        constexpr int NC = (NT == 2) ? 16 : 8;                 // windows (TMEM columns) per thread
        const int col0 = (NT == 2) ? 0 : wg * 8;
        const int j = (warp & 3) * 32 + lane;                  // hidden unit == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tile * 48 + col0);
        const float bhn = b_hn[dir * H + j];
        const int64_t wbase = win0 + tile * RT_N + col0;       // first window of this thread's columns
        uint8_t *hplane = smem + L::h_off + tile * 2 * RT_HPLANE;
        const uint32_t h_elem_off = (uint32_t)((j >> 3) * (RT_N * 16) + (j & 7) * 2);   // + n*16
        const int kcol = dir * H + j;
        const int64_t gcol = (int64_t)dir * G3 + j;

        float hprev[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) hprev[c] = 0.f;
        float gnext[3][NC];
        {
            const int64_t t = dir ? (T - 1) : 0;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const bool ok = (wbase + c) < B;
                const float *row = gi + ((ok ? (wbase + c) : 0) * T + t) * GI_COLS + gcol;
#pragma unroll
                for (int g = 0; g < 3; ++g) gnext[g][c] = ok ? ldg_stream(row + g * H) : 0.f;
            }
        }
        // h_{-1} = 0 is already in smem: publish it
        tc_fence_before_sync();
        mbar_arrive(&h_ready[tile]);

        for (int64_t step = 0; step < T; ++step) {
            const int64_t t = dir ? (T - 1 - step) : step;
            const int64_t tn = dir ? (t - 1) : (t + 1);
            const bool more = step + 1 < T;
            mbar_wait(&acc_ready[tile], (uint32_t)(step & 1));
            tc_fence_after_sync();
#pragma unroll
            for (int c8 = 0; c8 < NC; c8 += 8) {
                uint32_t ar[8], az[8], an[8];
                tmem_ld_x8(t_lane + 0 * 16 + c8, ar);
                tmem_ld_x8(t_lane + 1 * 16 + c8, az);
                tmem_ld_x8(t_lane + 2 * 16 + c8, an);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = c8 + i;
                    const float r = sigmoid_fast(gnext[0][c] + __uint_as_float(ar[i]));
                    const float z = sigmoid_fast(gnext[1][c] + __uint_as_float(az[i]));
                    const float nn = tanh_fast(gnext[2][c] + r * (__uint_as_float(an[i]) + bhn));
                    const float h = fmaf(hprev[c] - nn, z, nn);   // (hx - n) * z + n, as ATen's gru_cell
                    hprev[c] = h;
                    __half hi, lo;
                    split_f16(h, hi, lo);
                    const int n = col0 + c;                       // row of the B operand tile
                    *reinterpret_cast<__half *>(hplane + h_elem_off + n * 16) = hi;
                    *reinterpret_cast<__half *>(hplane + RT_HPLANE + h_elem_off + n * 16) = lo;
                    const int64_t w = wbase + c;
                    const bool ok = w < B;
                    if (ok) {
                        const int64_t p = w * T + t;
                        if (OUT_TILES) {
                            __half *tb = reinterpret_cast<__half *>(reinterpret_cast<uint8_t *>(h_out) +
                                                                    (p / XT_ROWS) * (int64_t)XT_TILE_BYTES);
                            const int64_t off = (int64_t)(kcol >> 3) * (XT_ROWS * 8) + (p % XT_ROWS) * 8 + (kcol & 7);
                            tb[off] = hi;
                            tb[XT_PLANE_BYTES / 2 + off] = lo;
                        } else {
                            reinterpret_cast<float *>(h_out)[p * H2 + kcol] = h;
                        }
                    }
                    // software pipeline: this column's pre-activations of the NEXT step reuse the same
                    // registers; the loads complete under the next step's MMA
                    if (more && ok) {
                        const float *row = gi + (w * T + tn) * GI_COLS + gcol;
#pragma unroll
                        for (int g = 0; g < 3; ++g) gnext[g][c] = ldg_stream(row + g * H);
                    }
                }
            }
            fence_proxy_async_smem();     // h tile writes -> visible to the MMA's async-proxy reads
            tc_fence_before_sync();       // order our tcgen05.ld before the next MMA overwrites the accumulators
            mbar_arrive(&h_ready[tile]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, RT_TMEM_COLS);
    }
}

cudaError_t launch_rec_tc(const float *gi, const __half *w_hh_tc, const float *b_hn, void *h_out, int out_tiles,
                          int64_t B, int64_t T, int sm_count, cudaStream_t s) {
    if (B == 0 || T == 0) return cudaSuccess;
    const int64_t tiles = (B + RT_N - 1) / RT_N;
    // ping-pong (2 tiles per CTA) only pays once there are more tiles than SMs to run them one per CTA
    const bool two = tiles * NDIR > (int64_t)sm_count;
    cudaError_t e;
#define MDK_LAUNCH_REC(NTV, OT)                                                                              \
    do {                                                                                                     \
        auto kern = rec_tc_kernel<NTV, OT>;                                                                  \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, RecSmem<NTV>::total);    \
        if (e != cudaSuccess) return e;                                                                      \
        dim3 grid((unsigned)((tiles + NTV - 1) / NTV), NDIR);                                                \
        kern<<<grid, RT_THREADS, RecSmem<NTV>::total, s>>>(gi, w_hh_tc, b_hn, h_out, B, T);                  \
    } while (0)
    if (two) { if (out_tiles) MDK_LAUNCH_REC(2, true); else MDK_LAUNCH_REC(2, false); }
    else     { if (out_tiles) MDK_LAUNCH_REC(1, true); else MDK_LAUNCH_REC(1, false); }
#undef MDK_LAUNCH_REC
    return cudaGetLastError();
}

// =====================================================================================================
// Layer-1 input projection on tensor cores:  gi[p][blk*128 + j] = sum_k W_ih[blk*128 + j][k] * x[p][k] + bias
// Persistent, warp-specialised: grid = 6 weight blocks x CT CTAs; each CTA keeps its 128x256 weight block
// (hi+lo, 128 KiB) resident and streams 128-position activation tiles (written by the layer-0 recurrent kernel
// directly in operand layout) through a 3-stage ring of 64-wide K slices with bulk async copies.
//   warp 0 : producer (cp.async.bulk -> smem, mbarrier complete_tx)
//   warp 1 : MMA issuer, TMEM owner (2 x 128-column accumulators, double buffered)
//   warps 2-5 : epilogue (TMEM -> registers -> + bias -> coalesced fp32 stores)
// =====================================================================================================
constexpr int GT_THREADS = 192;
constexpr int GT_STAGES = 3;
constexpr int GT_KSLICE = 64;                                   // K per stage
constexpr int GT_SLICE_BYTES = XT_ROWS * GT_KSLICE * 2;         // one plane of one slice: 16 KiB
constexpr int GT_STAGE_BYTES = 2 * GT_SLICE_BYTES;              // hi + lo
constexpr int GT_A_BYTES = 2 * H * H2 * 2;                      // 131 072
constexpr int GT_BAR_OFF = GT_A_BYTES + GT_STAGES * GT_STAGE_BYTES;
constexpr int GT_SMEM = GT_BAR_OFF + 128;
constexpr uint32_t GT_TMEM_COLS = 256;

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const uint8_t *__restrict__ x_tiles, const __half *__restrict__ w_in_tc,
               const float *__restrict__ bias, float *__restrict__ gi, int64_t P, int64_t ntiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + GT_BAR_OFF);
    uint64_t *empty = full + GT_STAGES;
    uint64_t *acc_full = empty + GT_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int blk = blockIdx.y;                 // weight block: dir*3 + gate
    const int64_t tile0 = blockIdx.x;
    const int64_t tstride = gridDim.x;

    {
        const int4 *src = reinterpret_cast<const int4 *>(reinterpret_cast<const uint8_t *>(w_in_tc) +
                                                         (size_t)blk * GT_A_BYTES);
        int4 *dst = reinterpret_cast<int4 *>(smem);
        for (int i = tid; i < GT_A_BYTES / 16; i += GT_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int i = 0; i < GT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, GT_TMEM_COLS);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = tile0; tile < ntiles; tile += tstride) {
                const uint8_t *src = x_tiles + tile * (int64_t)XT_TILE_BYTES;
                for (int s = 0; s < XT_K / GT_KSLICE; ++s, ++it) {
                    const uint32_t stage = it % GT_STAGES;
                    mbar_wait(&empty[stage], ((it / GT_STAGES) & 1) ^ 1);
                    uint8_t *dst = smem + GT_A_BYTES + stage * GT_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], GT_STAGE_BYTES);
                    bulk_g2s(dst, src + (size_t)s * GT_SLICE_BYTES, GT_SLICE_BYTES, &full[stage]);
                    bulk_g2s(dst + GT_SLICE_BYTES, src + XT_PLANE_BYTES + (size_t)s * GT_SLICE_BYTES,
                             GT_SLICE_BYTES, &full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = make_idesc_f16(128, XT_ROWS);
        const uint32_t a_addr = smem_u32(smem);
        const uint32_t b_addr = smem_u32(smem + GT_A_BYTES);
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = tile0; tile < ntiles; tile += tstride, ++tcount) {
            const uint32_t as = tcount & 1;
            mbar_wait(&acc_empty[as], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + as * XT_ROWS;
            for (int s = 0; s < XT_K / GT_KSLICE; ++s, ++it) {
                const uint32_t stage = it % GT_STAGES;
                mbar_wait(&full[stage], (it / GT_STAGES) & 1);
                tc_fence_after_sync();
                if (lane == 0) {
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        const int pa = (prod == 2) ? 1 : 0;   // W part
                        const int pb = (prod == 1) ? 1 : 0;   // x part
                        const uint32_t a0 = a_addr + pa * (H * H2 * 2) + s * (GT_KSLICE / 8) * (H * 16);
                        const uint32_t b0 = b_addr + stage * GT_STAGE_BYTES + pb * GT_SLICE_BYTES;
#pragma unroll
                        for (int ks = 0; ks < GT_KSLICE / 16; ++ks) {
                            const uint64_t ad = make_smem_desc(a0 + ks * 2 * (H * 16), H * 16, 128);
                            const uint64_t bd = make_smem_desc(b0 + ks * 2 * (XT_ROWS * 16), XT_ROWS * 16, 128);
                            umma_f16(d, ad, bd, idesc, (s | prod | ks) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (s == XT_K / GT_KSLICE - 1) umma_commit(&acc_full[as]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int j = q * 32 + lane;
        const float bj = bias[blk * H + j];
        uint32_t tcount = 0;
        for (int64_t tile = tile0; tile < ntiles; tile += tstride, ++tcount) {
            const uint32_t as = tcount & 1;
            mbar_wait(&acc_full[as], (tcount >> 1) & 1);
            tc_fence_after_sync();
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + as * XT_ROWS;
            float *out = gi + (tile * XT_ROWS) * (int64_t)GI_COLS + blk * H + j;
            const int64_t prem = P - tile * XT_ROWS;
#pragma unroll 1
            for (int c32 = 0; c32 < XT_ROWS; c32 += 32) {
                uint32_t v[32];
                tmem_ld_x32(t_lane + c32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (c32 + i < prem) out[(int64_t)(c32 + i) * GI_COLS] = __uint_as_float(v[i]) + bj;
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&acc_empty[as]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, GT_TMEM_COLS);
    }
}

cudaError_t launch_gemm_tc(const void *x_tiles, const __half *w_in_tc, const float *bias, float *gi, int64_t P,
                           int sm_count, cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    const int64_t ntiles = (P + XT_ROWS - 1) / XT_ROWS;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int64_t ct = sm_count / 6;
    if (ct < 1) ct = 1;
    if (ct > ntiles) ct = ntiles;
    dim3 grid((unsigned)ct, 6);
    gemm_tc_kernel<<<grid, GT_THREADS, GT_SMEM, s>>>(reinterpret_cast<const uint8_t *>(x_tiles), w_in_tc, bias, gi,
                                                     P, ntiles);
    return cudaGetLastError();
}

// =====================================================================================================
// Self test of the UMMA building block: D[128][N] = A[128][K] . B[N][K]^T, fp16 hi/lo split, one CTA.
// variant 0 = production descriptors; 1 = LBO/SBO swapped; 2 = descriptor version bits cleared.
// =====================================================================================================
__global__ void __launch_bounds__(128, 1)
selftest_kernel(const float *__restrict__ A, const float *__restrict__ Bm, float *__restrict__ D, int N, int K,
                int variant) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int a_plane = 128 * K * 2, b_plane = N * K * 2;
    uint8_t *sa = smem, *sb = smem + 2 * a_plane;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sb + 2 * b_plane);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        __half hi, lo;
        split_f16(A[i], hi, lo);
        const int off = (k / 8) * (128 * 16) + r * 16 + (k % 8) * 2;
        *reinterpret_cast<__half *>(sa + off) = hi;
        *reinterpret_cast<__half *>(sa + a_plane + off) = lo;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        __half hi, lo;
        split_f16(Bm[i], hi, lo);
        const int off = (k / 8) * (N * 16) + r * 16 + (k % 8) * 2;
        *reinterpret_cast<__half *>(sb + off) = hi;
        *reinterpret_cast<__half *>(sb + b_plane + off) = lo;
    }
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(tmem_slot, 128); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        uint32_t acc = 0;
        for (int prod = 0; prod < 3; ++prod) {
            const int pa = (prod == 2), pb = (prod == 1);
            for (int ks = 0; ks < K / 16; ++ks) {
                uint32_t a_lbo = 128 * 16, a_sbo = 128, b_lbo = N * 16, b_sbo = 128;
                if (variant == 1) { uint32_t t = a_lbo; a_lbo = a_sbo; a_sbo = t; t = b_lbo; b_lbo = b_sbo; b_sbo = t; }
                uint64_t ad = make_smem_desc(smem_u32(sa + pa * a_plane) + ks * 2 * 128 * 16, a_lbo, a_sbo);
                uint64_t bd = make_smem_desc(smem_u32(sb + pb * b_plane) + ks * 2 * N * 16, b_lbo, b_sbo);
                if (variant == 2) { ad &= ~(3ull << 46); bd &= ~(3ull << 46); }
                umma_f16(tmem_base, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    tc_fence_after_sync();
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    for (int c8 = 0; c8 < N; c8 += 8) {
        uint32_t v[8];
        tmem_ld_x8(t_lane + c8, v);
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) D[row * N + c8 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 128); }
}

int selftest_umma(int device, const float *A, const float *B, float *D, int N, int K, int variant) {
    MDK_REQUIRE(N >= 16 && N <= 128 && N % 16 == 0, MDK_ERR_ARG, "selftest_umma: N must be a multiple of 16 in [16,128]");
    MDK_REQUIRE(K >= 16 && K <= 256 && K % 16 == 0, MDK_ERR_ARG, "selftest_umma: K must be a multiple of 16 in [16,256]");
    MDK_CUDA(cudaSetDevice(device));
    float *dA = nullptr, *dB = nullptr, *dD = nullptr;
    MDK_CUDA(cudaMalloc(&dA, sizeof(float) * 128 * K));
    MDK_CUDA(cudaMalloc(&dB, sizeof(float) * N * K));
    MDK_CUDA(cudaMalloc(&dD, sizeof(float) * 128 * N));
    MDK_CUDA(cudaMemcpy(dA, A, sizeof(float) * 128 * K, cudaMemcpyHostToDevice));
    MDK_CUDA(cudaMemcpy(dB, B, sizeof(float) * N * K, cudaMemcpyHostToDevice));
    const int smem = 2 * 128 * K * 2 + 2 * N * K * 2 + 64;
    MDK_REQUIRE(smem <= 227 * 1024, MDK_ERR_ARG, "selftest_umma: N*K too large for shared memory");
    MDK_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    selftest_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, variant);
    MDK_CUDA(cudaGetLastError());
    MDK_CUDA(cudaDeviceSynchronize());
    MDK_CUDA(cudaMemcpy(D, dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return MDK_OK;
}

}  // namespace mdk
