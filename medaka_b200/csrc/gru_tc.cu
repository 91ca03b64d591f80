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
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"
#include "rec_common.cuh"

namespace mdk {

// =====================================================================================================
// Recurrent kernel.  One CTA = NT tiles of 16 windows of one direction, for the whole sequence.
//   warps 0-15  : gate warps (TMEM -> registers -> gate math -> next h into smem + global output); warp w reads
//                 TMEM lane quarter w%4 and the column group w/4
//   warps 16-18 : service warps (warp 16 also owns the TMEM allocation)
// NT == 1: all four gate-warp groups share tile 0 (4 windows per thread): used when there are too few windows to give
//          every SM two tiles - the BASELINE 10 Mb workload.  Warp 16 = the MMA issuer (r, z, n in that order, one commit
//          each, then - LOGITS - the linear head's MMAs), warp 17 = relay (sole waiter on the commit mbarriers, releases
//          the gate warps through named barriers), warp 18 = aux (gi staging by bulk copy, L2 prefetch, bulk copy-out
//          of the h tile).  See the issuer section for the protocol and DESIGN.md section 6 for where the cycles go.
// NT == 2: gate-warp groups 0,1 own tile 0 and groups 2,3 tile 1 (8 windows per thread); warps 16-18 issue one gate
//          block each; the MMAs of one tile overlap the gate math of the other (ping-pong).
//
// W_hh (fp16 hi+lo, 384 columns) lives in TENSOR MEMORY for the whole sequence and is the A operand of the
// tcgen05.mma ".ts" form: with N = 16 an SS-mode MMA re-reads a 4 KiB A tile from shared memory for 8 cycles of
// tensor work (measured in round 1a: tensor-core smem reads 41 % busy; profiles/r01a_*).
//
// FUSE_X (layer 0): the input projection W_ih . x_t (K = F <= 16) is done here as 9 extra MMAs per tile-step
// (x_t staged as a 16x16 fp16 hi/lo B tile by the gate warps one step ahead), so layer 0 needs no gi buffer at
// all: the separate projection kernel and its 6 KiB/position HBM round trip disappear, and the gate loop has no
// global loads.  The n gate needs W_in.x and W_hn.h apart, hence a 4th 16-column accumulator per tile.
// TMEM budget: 384 (W_hh) [+ 48 (W_ih, NT == 1 only; NT == 2 reads it from smem in SS mode)] + NT x 48/64 <= 512.
// =====================================================================================================
constexpr int RT_GATE_WARPS = 16;                        // 4 per scheduler: the gate phase is latency-bound
constexpr int RT_MMA_WARPS = 3;                          // service warps: NT == 2 one issuer per gate block; NT == 1 issuer, relay, aux
constexpr int RT_THREADS = 32 * (RT_GATE_WARPS + RT_MMA_WARPS);
// Hand-off protocol.  MDK_REC_NB = 1 (default): the gate warps -> issuers hand-off ("h tile written") is a NAMED hardware
// barrier (bar.arrive by the 16 gate warps, bar.sync by the issuers): the cycle trace (profiles/r01e) shows the issuers
// released ~70 cycles after the last gate warp arrives, against ~190 with 512 per-thread mbarrier arrivals.  The
// accumulator hand-off stays an mbarrier (tcgen05.commit needs one) polled by the gate warps directly: relaying it
// through an issuer warp and a named barrier measured slower.  MDK_REC_NB = 0 keeps the all-mbarrier protocol (A/B).
#ifndef MDK_REC_NB
#define MDK_REC_NB 1
#endif
constexpr bool RT_NB = MDK_REC_NB != 0;
#ifndef MDK_REC_POLL1
#define MDK_REC_POLL1 1
#endif
#if MDK_REC_POLL1
#define GATE_WAIT(bar, par) mbar_wait_warp(bar, par)
#else
#define GATE_WAIT(bar, par) mbar_wait(bar, par)
#endif
constexpr int RT_BAR_H = 1;        // ids 1, 2: h tile of tile 0 / 1 written (gate warps arrive, issuers sync)
constexpr int RT_BAR_R = 3;        // NT == 1: r / z / n accumulators complete (relay warp arrives, gate warps sync)
constexpr int RT_BAR_Z = 4;
constexpr int RT_BAR_N = 5;
constexpr int RT_WX_BLOCK = H * 16 * 2;                  // one (part, gate) block of W_ih in smem: [kg 2][row 128][8] = 4 KiB

template <int NT, bool FUSE_X, bool LOGITS = false>
struct RecCfg {
    static constexpr bool wx_tmem = FUSE_X && NT == 1;
    // accumulator columns per tile: r, z, n (16 each) [+ W_in.x of the n gate]
    static constexpr int acc_per_tile = FUSE_X ? 64 : 48;
    static constexpr uint32_t wx_col0 = RT_WT_COLS;
    static constexpr uint32_t acc_col0 = RT_WT_COLS + (wx_tmem ? RT_WX_COLS : 0);
    static_assert(acc_col0 + NT * acc_per_tile <= 512, "TMEM budget");
    static constexpr int h_off = 0;                                        // [NT][2 planes][RT_HPLANE]
    static constexpr int x_off = ((NT * 2 * RT_HPLANE + 127) / 128) * 128; // [NT][2 bufs][RT_XBUF]
    static constexpr int wx_off = x_off + ((NT * 2 * RT_XBUF + 127) / 128) * 128;   // [6][RT_WX_BLOCK] (NT == 2)
    // NT == 1 splits the accumulator hand-off: the r/z blocks are committed (and their sigmoids start) while the
    // n-gate MMAs still run; with NT == 2 the other tile already fills that time
    static constexpr bool split = (NT == 1);
    static constexpr int bar_off = wx_off + 6 * RT_WX_BLOCK;               // acc_ready[NT], h_ready[NT], acc_n[NT], rz_issued[NT]
    static constexpr int tmem_off = bar_off + 4 * NT * 8;
    // NT == 1 reading gi (layer 1): the 24 KiB block of (tile-step, direction) is staged in shared memory by a bulk
    // async copy two steps ahead (3 buffers), so the gate warps read their pre-activations with three LDS.128 instead of
    // streaming 24 KiB per step through the LSU global path, which cost them 400-900 cycles per step (cycle trace)
    static constexpr bool gi_smem = !FUSE_X && NT == 1;
    static constexpr int GI_BUFS = 3;
    static constexpr int GI_BLOCK = (GI_TS_FLOATS / 2) * 4;                // 24 576 bytes
    static constexpr int gibar_off = tmem_off + 16;                        // gi_full[GI_BUFS]
    static constexpr int gi_off = ((gibar_off + GI_BUFS * 8 + 127) / 128) * 128;
    static constexpr int gi_end = gi_off + (gi_smem ? GI_BUFS * GI_BLOCK : 0);
    // the CTA owns all 512 TMEM columns of its SM, so a second co-resident CTA could only spin in tcgen05.alloc;
    // ask for > half of the shared memory to keep residency at one CTA per SM.
    // LOGITS (layer 1, NT == 1): the linear head runs as 24 extra M64 N16 K16 MMAs per step on the h tile.  They execute
    // while the gate warps are already producing the next h, so the tile is DOUBLE BUFFERED (HB = distance between the
    // buffers; the tile read by step t's MMAs is only overwritten in the tail of step t + 1, after that step's n commit,
    // which the in-order tensor pipe completes after these MMAs).  W_lin (fp16 hi/lo, 64 rows, 5 used) is a shared-memory
    // A operand image [plane][k-group 16][row 64][8 halfs].
    static constexpr int h2_off = ((gi_end + 127) / 128) * 128;
    static constexpr int HB = LOGITS ? h2_off - h_off : 0;
    static constexpr int WL_PLANE = 16 * 64 * 16;                          // 16 KiB
    static constexpr int wl_off = h2_off + (LOGITS ? ((NT * 2 * RT_HPLANE + 127) / 128) * 128 : 0);
    static constexpr int logbar_off = wl_off + (LOGITS ? 2 * WL_PLANE : 0);
    static constexpr uint32_t log_col = acc_col0 + NT * acc_per_tile;      // 16 accumulator columns of the logits MMAs
    static constexpr uint32_t wl_col = log_col + 16;                       // W_lin hi plane as a TMEM A operand (64 columns)
    static_assert(!LOGITS || wl_col + H / 2 <= 512, "TMEM budget");
    static constexpr int end_ = logbar_off + 16;
    static constexpr int total = end_ > 120 * 1024 ? end_ : 120 * 1024;
    static_assert(total <= 227 * 1024, "smem budget");
};


// Diagnostics (TRACE instantiations only, selected by mdk_debug_rec_trace): CTA (0,0) stamps %clock64 at the hand-off
// points of time steps [RT_TRACE_STEP0, +RT_TRACE_STEPS) into trace[step][slot]; slots are listed in tools/diag.py.
constexpr int RT_TRACE_STEP0 = 512, RT_TRACE_STEPS = 16, RT_TRACE_SLOTS = 40;
#define REC_STAMP(slot)                                       \
    do {                                                      \
        if (TRACE && tr) tr[slot] = (unsigned long long)clock64(); \
    } while (0)

template <int NT, bool OUT_TILES, bool FUSE_X, bool TRACE = false, bool LOGITS = false>
__global__ void __launch_bounds__(RT_THREADS, 1)
rec_tc_kernel(const float *__restrict__ gi, RecX xin, const __half *__restrict__ w_hh,
              const float *__restrict__ b_hn, void *__restrict__ h_out, int64_t B, int64_t T,
              unsigned long long *__restrict__ trace, const __half *__restrict__ lin_w_tc, float *__restrict__ plog,
              uint32_t prod_mask) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = RecCfg<NT, FUSE_X, LOGITS>;
    static_assert(!LOGITS || (NT == 1 && !FUSE_X && !OUT_TILES), "fused logits: layer 1, one tile per CTA");
    uint64_t *log_bar = reinterpret_cast<uint64_t *>(smem + L::logbar_off);
    // NT == 1 writing operand tiles (layer 0): the h tile the gate warps publish in shared memory already IS the tile
    // image the projection GEMM wants (per k-group 16 rows x 16 B contiguous), so warp 18 copies it out with 32 bulk
    // async copies per step instead of 8 two-byte global stores per gate thread.
    constexpr bool BULK_OUT = OUT_TILES && NT == 1;
    constexpr int NB_N_COUNT = 32 * RT_GATE_WARPS + (BULK_OUT ? 64 : 32);   // BAR_N: gate warps + relay (+ copy-out warp)
    constexpr bool GI_SMEM = L::gi_smem;
    constexpr int NB_R_COUNT = 32 * RT_GATE_WARPS + (GI_SMEM ? 64 : 32);    // BAR_R: gate warps + relay (+ gi staging warp)
    uint64_t *gi_full = reinterpret_cast<uint64_t *>(smem + L::gibar_off);
    uint64_t *acc_ready = reinterpret_cast<uint64_t *>(smem + L::bar_off);   // split: r and z blocks only
    uint64_t *h_ready = acc_ready + NT;
    uint64_t *acc_n = h_ready + NT;          // split: n block(s)
    uint64_t *rz_issued = acc_n + NT;        // split: r and z issuers have queued their MMAs
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::tmem_off);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int dir = blockIdx.y;
    const int64_t win0 = (int64_t)blockIdx.x * (RT_N * NT);

    // ---- prologue: zero h / x tiles, barriers, TMEM, weights ----
    {
        int4 *z = reinterpret_cast<int4 *>(smem);
        for (int i = tid; i < L::wx_off / 16; i += RT_THREADS) z[i] = make_int4(0, 0, 0, 0);
    }
    if (FUSE_X && !L::wx_tmem) {
        // W_ih -> smem operand image [part*3+gate][kg 2][row 128][8 halfs]
        const __half *src = xin.w_x + (size_t)dir * 6 * H * 16;
        __half *dst = reinterpret_cast<__half *>(smem + L::wx_off);
        for (int i = tid; i < 6 * H * 16; i += RT_THREADS) {
            const int pg = i / (H * 16), r = (i / 16) % H, k = i % 16;
            dst[pg * (RT_WX_BLOCK / 2) + (k >> 3) * (H * 8) + r * 8 + (k & 7)] = src[i];
        }
    }
    if (LOGITS) {   // this direction's half of W_lin, already in operand layout
        const int4 *src = reinterpret_cast<const int4 *>(lin_w_tc + (size_t)dir * 2 * (L::WL_PLANE / 2));
        int4 *dst = reinterpret_cast<int4 *>(smem + L::wl_off);
        for (int i = tid; i < 2 * L::WL_PLANE / 16; i += RT_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int i = 0; i < NT; ++i) {
            // split (NT == 1): acc_ready / rz_issued / acc_n = commit of the r / z / n block, one arrival each
            mbar_init(&acc_ready[i], L::split ? 1 : RT_MMA_WARPS);
            mbar_init(&h_ready[i], 32 * RT_GATE_WARPS / NT);
            mbar_init(&acc_n[i], 1);
            mbar_init(&rz_issued[i], 1);
        }
        for (int i = 0; i < L::GI_BUFS; ++i) mbar_init(&gi_full[i], 1);
        if (LOGITS) mbar_init(log_bar, 1);
        fence_mbar_init();
    }
    if (warp == RT_GATE_WARPS) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0u) {   // a 512-column allocation starts at column 0; the MMA issuers rely on literal addresses
        if (tid == 0) printf("mdk: unexpected TMEM base %u for a 512-column allocation\n", tmem_base);
        __trap();
    }

    // weights (row-major fp16 hi/lo) -> TMEM: lane j, 8 columns per K=16 chunk, cell = (k even | k odd << 16)
    if (warp < 4) {
        const int jrow = warp * 32 + lane;
        const uint32_t t_w = (uint32_t)(warp * 32) << 16;
        for (int pg = 0; pg < 6; ++pg) {   // pg = part*3 + gate
            const uint4 *src = reinterpret_cast<const uint4 *>(w_hh + (((size_t)dir * 6 + pg) * H + jrow) * H);
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                const uint4 lo4 = src[2 * ks], hi4 = src[2 * ks + 1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + (uint32_t)((pg * 8 + ks) * 8), v);
            }
            if (LOGITS && pg == 0) {
                // W_lin hi plane (rows >= 5 zero), row-major copy behind the two shared-memory images
                const uint4 *sl = reinterpret_cast<const uint4 *>(lin_w_tc + (size_t)NDIR * 2 * (L::WL_PLANE / 2) +
                                                                  ((size_t)dir * H + jrow) * H);
#pragma unroll
                for (int ks = 0; ks < H / 16; ++ks) {
                    const uint4 lo4 = sl[2 * ks], hi4 = sl[2 * ks + 1];
                    const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                    tmem_st_x8(t_w + L::wl_col + (uint32_t)(ks * 8), v);
                }
            }
            if (L::wx_tmem) {
                const uint4 *sx = reinterpret_cast<const uint4 *>(xin.w_x + (((size_t)dir * 6 + pg) * H + jrow) * 16);
                const uint4 lo4 = sx[0], hi4 = sx[1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + L::wx_col0 + (uint32_t)(pg * 8), v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (warp >= RT_GATE_WARPS) {
        // ================= MMA issuers =================
        // Every operand must sit in UNIFORM registers, otherwise the compiler wraps each tcgen05.mma in an
        // R2UR + elect waterfall loop (~50 issue cycles per MMA, measured).  So: TMEM addresses are literals,
        // shared-memory descriptors derive from the constant dynamic-smem base, and the issue is predicated by
        // elect.sync, not by `lane == 0`.
        //
        // NT == 2: warp 16 + g issues gate block g of both tiles and all three commit to acc_ready[tile].
        // NT == 1: ONE issuer (warp 16) queues r, z, n in that order with a commit after each block, so the r
        // accumulators complete after 24 of the 72 MMAs and sigmoid(r), then sigmoid(z), run under the rest of the MMA
        // phase.  Warp 17 is the RELAY: it alone waits on the three commit barriers and releases the 16 gate warps
        // through named barriers - an idle single waiter sees a commit ~100 cycles after the MMAs finish, while 16 warps
        // blocked in mbarrier.try_wait were released another 100-250 cycles later, and even an already-completed
        // try_wait cost 75-160 cycles on the critical path (cycle trace, profiles/r01e_*).  Warp 18 prefetches.
        const int g = warp - RT_GATE_WARPS;
        const uint32_t idesc = make_idesc_f16(128, RT_N);
        const uint64_t b_desc0 = make_smem_desc(smem_u32(smem + L::h_off), RT_KG, 128);
        const uint64_t x_desc0 = make_smem_desc(smem_u32(smem + L::x_off), RT_KG, 128);
        const uint64_t wx_desc0 = make_smem_desc(smem_u32(smem + L::wx_off), H * 16, 128);
        // W_hh[gate] (K steps [ks0, ks1)) . h of `tile`, three fp16 products, into accumulator columns d
        uint64_t b_cur = b_desc0;      // descriptor of the h tile buffer this step's MMAs read (LOGITS: alternates)
        auto issue_h = [&](uint32_t d, int gate, int ks0, int ks1, int tile, bool fresh) {
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                if (prod && !(prod_mask & (1u << prod))) continue;   // precision experiments: drop a correction product
                const int pa = (prod == 2) ? 1 : 0;   // W part: hi, hi, lo
                const int pb = (prod == 1) ? 1 : 0;   // activation part: hi, lo, hi
#pragma unroll
                for (int ks = ks0; ks < ks1; ++ks) {
                    const uint64_t bd = b_cur + (uint64_t)(((tile * 2 + pb) * RT_HPLANE + ks * 2 * RT_KG) >> 4);
                    umma_f16_ts(d, (uint32_t)(((pa * 3 + gate) * 8 + ks) * 8), bd, idesc,
                                (fresh && prod == 0 && ks == ks0) ? 0u : 1u);
                }
            }
        };
        // logits of the position whose h is in the current tile: W_lin (M = 64 rows, 5 used; A from shared memory) . h
        const uint64_t wl_desc0 = make_smem_desc(smem_u32(smem + L::wl_off), 64 * 16, 128);
        const uint32_t idesc64 = make_idesc_f16(64, RT_N);
        auto issue_logits = [&]() {
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                if (prod && !(prod_mask & (1u << prod))) continue;
                const int pa = (prod == 2) ? 1 : 0;
                const int pb = (prod == 1) ? 1 : 0;
#pragma unroll
                for (int ks = 0; ks < H / 16; ++ks) {
                    const uint64_t bd = b_cur + (uint64_t)((pb * RT_HPLANE + ks * 2 * RT_KG) >> 4);
                    if (pa == 0) {
                        // W_lin hi from tensor memory (M = 128, ~10 cycles, no shared-memory A traffic)
                        umma_f16_ts(L::log_col, L::wl_col + (uint32_t)(ks * 8), bd, idesc, (prod | ks) ? 1u : 0u);
                    } else {
                        // W_lin lo from shared memory (M = 64: rows 0..4 land in the same TMEM lanes 0..4)
                        const uint64_t ad = wl_desc0 + (uint64_t)((L::WL_PLANE + ks * 2 * (64 * 16)) >> 4);
                        umma_f16(L::log_col, ad, bd, idesc64, 1u);
                    }
                }
            }
            umma_commit(log_bar);
        };
        // TMEM -> plog[dir][tile-step][class][16 windows]: lanes 0..4 of this warp (TMEM lanes 0..4 = classes) hold one row
        auto store_logits = [&](int64_t t_idx) {
            uint32_t v[16];
            tc_fence_after_sync();
            tmem_ld_x16(L::log_col, v);
            tmem_ld_wait();
            if (lane < NCLS) {
                float4 *dst = reinterpret_cast<float4 *>(
                    plog + (((int64_t)dir * gridDim.x + blockIdx.x) * T + t_idx) * PLOG_TS_FLOATS + lane * WT);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            tc_fence_before_sync();
        };
        uint32_t log_par = 0;
        // + W_ih[gate] . x_t (fused layer-0 input projection)
        auto issue_x = [&](uint32_t d, int gate, int tile, uint32_t par, bool fresh) {
#pragma unroll
            for (int prod = 0; prod < 3; ++prod) {
                if (prod && !(prod_mask & (1u << prod))) continue;
                const int pa = (prod == 2) ? 1 : 0;
                const int pb = (prod == 1) ? 1 : 0;
                const uint64_t xd = x_desc0 + (uint64_t)(((tile * 2 + (int)par) * RT_XBUF + pb * RT_XPLANE) >> 4);
                const uint32_t acc = (fresh && prod == 0) ? 0u : 1u;
                if (L::wx_tmem) {
                    umma_f16_ts(d, L::wx_col0 + (uint32_t)((pa * 3 + gate) * 8), xd, idesc, acc);
                } else {
                    umma_f16(d, wx_desc0 + (uint64_t)(((pa * 3 + gate) * RT_WX_BLOCK) >> 4), xd, idesc, acc);
                }
            }
        };
        // L2 prefetch duty (one elected thread): feature rows X_PREFETCH_AHEAD.. steps ahead (fused layer 0) or the
        // pre-activation rows GI_PREFETCH_STEPS ahead (layer 1).  The gate warps' own register prefetch runs 1-2 steps
        // ahead, which covers an L2 hit but not a DRAM miss (measured: ~600 cycles per step exposed without this).
        auto prefetch = [&](int64_t step, int tile) {
            const int64_t wt = (int64_t)blockIdx.x * NT + tile;
            if (FUSE_X) {
                if ((step & (X_PREFETCH_EVERY - 1)) != 0) return;
                const int64_t s0 = step + X_PREFETCH_AHEAD;                    // first step covered
                const int64_t s1 = s0 + X_PREFETCH_EVERY <= T ? s0 + X_PREFETCH_EVERY : T;   // one past the last
                if (s0 >= T) return;
                const int64_t t_lo = dir ? (T - s1) : s0;                      // lowest time index of the span
                const int64_t nbytes = (s1 - s0) * xin.F * 4;
                for (int w = 0; w < WT; ++w) {
                    if (wt * WT + w >= B) break;
                    const uintptr_t a = reinterpret_cast<uintptr_t>(xin.feats + ((wt * WT + w) * T + t_lo) * xin.F);
                    const uintptr_t a0 = a & ~(uintptr_t)15;
                    bulk_prefetch_l2(reinterpret_cast<const void *>(a0),
                                     (uint32_t)(((a + nbytes - a0) + 15) & ~(uintptr_t)15));
                }
            } else {
                if (step + GI_PREFETCH_STEPS >= T || wt * WT >= B) return;
                const int64_t sp = step + GI_PREFETCH_STEPS;
                const int64_t t = dir ? (T - 1 - sp) : sp;
                // quad layout: the three gate blocks of (tile-step, direction) are one contiguous 24 KiB range
                // (issued as 2 KiB pieces: a single 24 KiB prefetch measured as if it had not been issued at all)
                const float *blk = gi + (wt * T + t) * GI_TS_FLOATS + (int64_t)dir * (GI_TS_FLOATS / 2);
#pragma unroll
                for (int i = 0; i < 12; ++i) bulk_prefetch_l2(blk + i * 512, 2048);
            }
        };
        constexpr int HCOUNT = 32 * RT_GATE_WARPS / NT + 32 * RT_MMA_WARPS;
#pragma unroll 1
        for (int64_t step = 0; step < T; ++step) {
            const uint32_t par = (uint32_t)(step & 1);
            unsigned long long *tr = nullptr;
            if (TRACE && trace && blockIdx.x == 0 && blockIdx.y == 0 && step >= RT_TRACE_STEP0 &&
                step < RT_TRACE_STEP0 + RT_TRACE_STEPS)
                tr = trace + (step - RT_TRACE_STEP0) * RT_TRACE_SLOTS;
            if (LOGITS) b_cur = b_desc0 + (uint64_t)((par * L::HB) >> 4);
#pragma unroll
            for (int tile = 0; tile < NT; ++tile) {
                if (RT_NB) {
                    if (tile == 0) named_bar_sync<RT_BAR_H, HCOUNT>(); else named_bar_sync<RT_BAR_H + 1, HCOUNT>();
                } else {
                    mbar_wait(&h_ready[tile], par);
                }
                tc_fence_after_sync();
                if (GI_SMEM && g == 2 && step > 0) {
                    // this step's gi block landed long ago (requested two steps back): release the gate warps' BAR_R
                    // before doing anything else (step 0: after the first blocks are requested, below)
                    mbar_wait(&gi_full[step % L::GI_BUFS], (uint32_t)((step / L::GI_BUFS) & 1));
                    named_bar_arrive<RT_BAR_R, NB_R_COUNT>();
                    if (TRACE && tr && lane == 0) tr[35] = (unsigned long long)clock64();
                }
                if (elect_one()) {
                    const uint32_t d0 = L::acc_col0 + (uint32_t)(tile * L::acc_per_tile);
                    if (L::split) {
                        if (g == 0) {
                            REC_STAMP(0);
                            issue_h(d0, 0, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0, 0, tile, par, false);
                            umma_commit(&acc_ready[tile]);
                            REC_STAMP(1);
                            issue_h(d0 + 16, 1, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0 + 16, 1, tile, par, false);
                            umma_commit(&rz_issued[tile]);
                            REC_STAMP(2);
                            issue_h(d0 + 32, 2, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0 + 48, 2, tile, par, true);
                            umma_commit(&acc_n[tile]);
                            REC_STAMP(3);
                            // the tile holds h of the previous step: its logits ride in the gate phase's shadow
                            if (LOGITS && step > 0) issue_logits();
                        } else if (g == 2) {
                            if (BULK_OUT && step > 0) {
                                // h_{step-1} (published through BAR_H) -> its rows of the GEMM operand tiles
                                const int64_t orow = ((int64_t)blockIdx.x * T + (dir ? (T - step) : (step - 1))) * WT;
                                uint8_t *dst = reinterpret_cast<uint8_t *>(h_out) + (orow >> 7) * (int64_t)XT_TILE_BYTES +
                                               (int64_t)(dir * (H / 8)) * (XT_ROWS * 16) + (orow & (XT_ROWS - 1)) * 16;
#pragma unroll
                                for (int plane = 0; plane < 2; ++plane)
#pragma unroll
                                    for (int kg = 0; kg < H / 8; ++kg)
                                        bulk_s2g(dst + plane * XT_PLANE_BYTES + kg * (XT_ROWS * 16),
                                                 smem + L::h_off + plane * RT_HPLANE + kg * RT_KG, WT * 16);
                                bulk_commit_group();
                            }
                            if (GI_SMEM) {
                                // stage the gi block of step + 2 into the buffer step - 1 used (free: every gate warp
                                // has passed BAR_H of this step, i.e. finished step - 1); steps 0 and 1 on the first pass
                                const int64_t s0 = step == 0 ? 0 : step + 2, s1 = step + 2;
                                for (int64_t sp = s0; sp <= s1 && sp < T; ++sp) {
                                    const int64_t t = dir ? (T - 1 - sp) : sp;
                                    const float *src = gi + ((int64_t)blockIdx.x * T + t) * GI_TS_FLOATS +
                                                       (int64_t)dir * (GI_TS_FLOATS / 2);
                                    const int b = (int)(sp % L::GI_BUFS);
                                    mbar_arrive_expect_tx(&gi_full[b], L::GI_BLOCK);
                                    bulk_g2s(smem + L::gi_off + b * L::GI_BLOCK, src, L::GI_BLOCK, &gi_full[b]);
                                }
                            }
                            REC_STAMP(32);
                            prefetch(step, tile);
                            REC_STAMP(33);
                            // the gate warps overwrite the tile after BAR_N: the copies must have read it by then
                            if (BULK_OUT && step > 0) bulk_wait_read_all();
                            REC_STAMP(34);
                        }
                    } else {
                        // one issuer per gate block; r and z accumulate the input projection onto the recurrent sum,
                        // n keeps W_in.x in its own columns
                        if (g == 0) {
                            issue_h(d0, 0, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0, 0, tile, par, false);
                        } else if (g == 1) {
                            issue_h(d0 + 16, 1, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0 + 16, 1, tile, par, false);
                        } else {
                            issue_h(d0 + 32, 2, 0, H / 16, tile, true);
                            if (FUSE_X) issue_x(d0 + 48, 2, tile, par, true);
                        }
                        umma_commit(&acc_ready[tile]);
                        if (g == 2) prefetch(step, tile);
                    }
                }
                __syncwarp();
                if (LOGITS && g == 0 && step > 0) {
                    mbar_wait(log_bar, log_par);
                    log_par ^= 1u;
                    store_logits(dir ? (T - step) : (step - 1));
                }
                if (BULK_OUT && g == 2) named_bar_arrive<RT_BAR_N, NB_N_COUNT>();
                if (GI_SMEM && g == 2 && step == 0) {
                    mbar_wait(&gi_full[0], 0u);
                    named_bar_arrive<RT_BAR_R, NB_R_COUNT>();
                }

                if (L::split && g == 1) {
                    constexpr int RC = 32 * RT_GATE_WARPS + 32;
                    mbar_wait(&acc_ready[tile], par);
                    tc_fence_after_sync();
                    tc_fence_before_sync();
                    named_bar_arrive<RT_BAR_R, NB_R_COUNT>();
                    if (TRACE && tr && lane == 0) tr[13] = (unsigned long long)clock64();
                    mbar_wait(&rz_issued[tile], par);
                    tc_fence_after_sync();
                    tc_fence_before_sync();
                    named_bar_arrive<RT_BAR_Z, RC>();
                    if (TRACE && tr && lane == 0) tr[14] = (unsigned long long)clock64();
                    mbar_wait(&acc_n[tile], par);
                    tc_fence_after_sync();
                    tc_fence_before_sync();
                    named_bar_arrive<RT_BAR_N, NB_N_COUNT>();
                    if (TRACE && tr && lane == 0) tr[15] = (unsigned long long)clock64();
                }
            }
        }
        if (LOGITS && g == 0) {
            // h of the last step: wait for every gate warp's final tile (named barrier id 6), then one more round
            named_bar_sync<6, 32 * RT_GATE_WARPS + 32>();
            tc_fence_after_sync();
            b_cur = b_desc0 + (uint64_t)((((uint32_t)T & 1u) * L::HB) >> 4);
            if (elect_one()) issue_logits();
            __syncwarp();
            mbar_wait(log_bar, log_par);
            store_logits(dir ? 0 : (T - 1));
        }
    } else {
        // ================= gate warps =================
        // All intermediates use the tile-interleaved row order (common.cuh): row(w, t) = ((w/16)*T + t)*16 + w%16,
        // so this thread's NC windows are NC consecutive rows: every address below is `step base + constant`.
        // Padding windows of a ragged last tile are ordinary rows (never copied out), hence no predicates.
        const int grp = warp >> 2;                             // column group 0..3
        const int tile = (NT == 2) ? (grp >> 1) : 0;
        constexpr int NC = (NT == 2) ? 8 : 4;                  // windows (TMEM columns) per thread
        constexpr int NP = NC / 2;                             // ... processed as packed fp32 pairs
        const int col0 = (NT == 2) ? (grp & 1) * 8 : grp * 4;
        const int j = (warp & 3) * 32 + lane;                  // hidden unit == TMEM lane
        const uint32_t t_lane = ((uint32_t)((warp & 3) * 32) << 16) + L::acc_col0 + (uint32_t)(tile * L::acc_per_tile + col0);
        const int64_t wtile = blockIdx.x * NT + tile;          // window tile (16 windows) of this warp group
        // NT == 2 with an odd number of window tiles: the last CTA's second tile does not exist - it still runs the
        // barrier protocol (on zeros) but must not touch global memory
        const bool tile_ok = wtile * WT < B;
        uint8_t *hrow = smem + L::h_off + tile * 2 * RT_HPLANE + (j >> 3) * RT_KG + (j & 7) * 2 + col0 * 16;
        const int kcol = dir * H + j;
        const int64_t row0 = (wtile * T) * WT + col0;          // row of (column 0, t = 0); row(t) = row0 + t*16
        const int64_t tstep = dir ? -(int64_t)WT : (int64_t)WT;   // rows per time step, signed by direction
        const int64_t t_first = dir ? (T - 1) : 0;

        // unfused path: pre-activations from the gi buffer, prefetched one step ahead into registers
        // (quad layout, common.cuh: this thread's 3 gates x NC windows are 3 x NC/4 16-byte loads)
        const float4 *gptr = FUSE_X ? nullptr
                                    : reinterpret_cast<const float4 *>(gi) + (wtile * T + t_first) * (GI_TS_FLOATS / 4) +
                                          (int64_t)((dir * 3) * 4 + (col0 >> 2)) * H + j;
        const int64_t gstep = dir ? -(int64_t)(GI_TS_FLOATS / 4) : (int64_t)(GI_TS_FLOATS / 4);
        F2 g2[3][NP];
        int gbuf = 0;                                          // GI_SMEM: staging buffer of the current step
        auto load_gi = [&]() {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt)
#pragma unroll
                for (int c = 0; c < NC / 4; ++c) {
                    const float4 v = ld_stream4(reinterpret_cast<const float *>(gptr + (gt * 4 + c) * H));
                    g2[gt][2 * c] = f2_make(v.x, v.y);
                    g2[gt][2 * c + 1] = f2_make(v.z, v.w);
                }
        };
        // fused path: folded biases, and this thread's share of the x_t staging (16 x F values per tile-step)
        const F2 bhn2 = f2_make(b_hn[dir * H + j], b_hn[dir * H + j]);
        F2 br2 = f2_make(0.f, 0.f), bz2 = br2, bn2 = br2;
        const int tix = (NT == 2) ? (tid & 255) : tid;         // thread index within the tile's gate warps
        const int xF = FUSE_X ? xin.F : 1;                     // 1 <= F <= 16 on the fused path
        const int xn = tix / xF, xf = tix - xn * xF;           // window row / feature of the staged value
        const bool xown = FUSE_X && tix < RT_N * xF;
        const bool xok = xown && (wtile * WT + xn) < B;
        const float *xsrc = nullptr;
        uint8_t *xdst = nullptr;
        float xreg = 0.f;
        if (FUSE_X) {
            br2 = f2_make(xin.bias[dir * G3 + j], xin.bias[dir * G3 + j]);
            bz2 = f2_make(xin.bias[dir * G3 + H + j], xin.bias[dir * G3 + H + j]);
            bn2 = f2_make(xin.bias[dir * G3 + 2 * H + j], xin.bias[dir * G3 + 2 * H + j]);
            if (xown) {
                xsrc = xin.feats + ((wtile * WT + xn) * T) * xin.F + xf;
                xdst = smem + L::x_off + tile * 2 * RT_XBUF + (xf >> 3) * RT_KG + xn * 16 + (xf & 7) * 2;
            }
        }
        // output pointers at (column 0, first time step); advanced by a constant every step
        float *o32 = reinterpret_cast<float *>(h_out) + (row0 + t_first * WT) * H2 + kcol;
        int64_t orow = row0 + t_first * WT;                    // OUT_TILES: row -> (tile, row in tile)
        __half *o16 = reinterpret_cast<__half *>(h_out) + (int64_t)(kcol >> 3) * (XT_ROWS * 8) + (kcol & 7);

        const F2 one2 = f2_make(1.0f, 1.0f), negone2 = f2_make(-1.0f, -1.0f);
        F2 hprev2[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) hprev2[q] = f2_make(0.f, 0.f);
        if (!FUSE_X) {
#pragma unroll
            for (int q = 0; q < NP; ++q)
#pragma unroll
                for (int gt = 0; gt < 3; ++gt) g2[gt][q] = f2_make(0.f, 0.f);
            if (tile_ok && !GI_SMEM) load_gi();
        } else if (xown) {
            // x of step 0 -> buffer 0; x of step 1 -> register
            const float x0 = xok ? xsrc[t_first * (int64_t)xin.F] : 0.f;
            __half hi, lo;
            split_f16(x0, hi, lo);
            *reinterpret_cast<__half *>(xdst) = hi;
            *reinterpret_cast<__half *>(xdst + RT_XPLANE) = lo;
            if (T > 1 && xok) xreg = xsrc[(dir ? (T - 2) : 1) * (int64_t)xin.F];
        }
        // running pointer to the feature value of step + 2 (loaded while step is in its epilogue)
        const float *xnext = (FUSE_X && xown) ? xsrc + (dir ? (T - 3) : 2) * (int64_t)xin.F : nullptr;
        const int64_t xadv = dir ? -(int64_t)xin.F : (int64_t)xin.F;
        // h_{-1} = 0 (and x_0) are in smem: publish
        constexpr int HCOUNT = 32 * RT_GATE_WARPS / NT + 32 * RT_MMA_WARPS;
        fence_proxy_async_smem();
        tc_fence_before_sync();
        if (RT_NB) {
            if (tile == 0) named_bar_arrive<RT_BAR_H, HCOUNT>(); else named_bar_arrive<RT_BAR_H + 1, HCOUNT>();
        } else {
            mbar_arrive(&h_ready[tile]);
        }

#pragma unroll 1
        for (int64_t step = 0; step < T; ++step) {
            const bool more = step + 1 < T;
            unsigned long long *tr = nullptr;
            if (TRACE && trace && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && step >= RT_TRACE_STEP0 &&
                step < RT_TRACE_STEP0 + RT_TRACE_STEPS)
                tr = trace + (step - RT_TRACE_STEP0) * RT_TRACE_SLOTS;
            uint32_t ar[NC], az[NC], an[NC], ax[NC];
            __half hh[NC], hl[NC];
            __half *tb = nullptr;
            F2 r2[NP], eb2[NP], nzb2[NP];    // r ; e^{-pre_z} ; -(1 + e^{-pre_z})
            if constexpr (L::split) {
                // three hand-offs per step, each a named barrier released by the relay warp: r, then z, then n
                constexpr int RC = 32 * RT_GATE_WARPS + 32;
                named_bar_sync<RT_BAR_R, NB_R_COUNT>();
                tc_fence_after_sync();
                tmem_ld_x4(t_lane + 0 * 16, ar);
                if (GI_SMEM) {
                    const float4 *gs = reinterpret_cast<const float4 *>(smem + L::gi_off + gbuf * L::GI_BLOCK) + (col0 >> 2) * H + j;
#pragma unroll
                    for (int gt = 0; gt < 3; ++gt) {
                        const float4 v = gs[gt * 4 * H];
                        g2[gt][0] = f2_make(v.x, v.y);
                        g2[gt][1] = f2_make(v.z, v.w);
                    }
                    gbuf = (gbuf == L::GI_BUFS - 1) ? 0 : gbuf + 1;
                }
                tmem_ld_wait();
                REC_STAMP(4);
                if (!FUSE_X) gptr += gstep;                    // block of the next time step
                if (OUT_TILES) tb = o16 + (orow >> 7) * (int64_t)(XT_TILE_BYTES / 2) + (orow & (XT_ROWS - 1)) * 8;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    // (weights and biases carry the -log2 e factor, common.cuh gate_scale: the accumulator is the exponent.)
                    // The reciprocal runs on the FMA pipe: with MUFU.RCP this phase was bound by the XU pipe (8 MUFU per
                    // thread, 256 cycles per SM sub-partition) and sits on the r -> z -> n gate chain.
                    const F2 accr = f2_make(__uint_as_float(ar[2 * q]), __uint_as_float(ar[2 * q + 1]));
                    float a0, a1;
                    f2_get(f2_add(FUSE_X ? br2 : g2[0][q], accr), a0, a1);
                    const F2 ea = f2_make(ex2_approx(fminf(a0, EXP_CLAMP)), ex2_approx(fminf(a1, EXP_CLAMP)));
                    r2[q] = rcp_neg_fma2(f2_fma(ea, negone2, negone2), one2);    // r = 1 / (1 + ea)
                }
                REC_STAMP(5);
                phase_fence(r2[0], r2[NP - 1]);
                named_bar_sync<RT_BAR_Z, RC>();
                tc_fence_after_sync();
                tmem_ld_x4(t_lane + 1 * 16, az);
                tmem_ld_wait();
                REC_STAMP(6);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    // z = 1 / (1 + eb) is never formed: its reciprocal is shared with the tanh below
                    const F2 accz = f2_make(__uint_as_float(az[2 * q]), __uint_as_float(az[2 * q + 1]));
                    float b0, b1;
                    f2_get(f2_add(FUSE_X ? bz2 : g2[1][q], accz), b0, b1);
                    eb2[q] = f2_make(ex2_approx(fminf(b0, EXP_CLAMP)), ex2_approx(fminf(b1, EXP_CLAMP)));
                    nzb2[q] = f2_fma(eb2[q], negone2, negone2);          // -(1 + eb)
                }
                REC_STAMP(7);
                phase_fence(nzb2[0], nzb2[NP - 1]);
                named_bar_sync<RT_BAR_N, NB_N_COUNT>();
                tc_fence_after_sync();
                tmem_ld_x4(t_lane + 2 * 16, an);
                if (FUSE_X) tmem_ld_x4(t_lane + 3 * 16, ax);
                tmem_ld_wait();
            } else {
                GATE_WAIT(&acc_ready[tile], (uint32_t)(step & 1));
                tc_fence_after_sync();
                if constexpr (NC == 8) {
                    tmem_ld_x8(t_lane + 0 * 16, ar);
                    tmem_ld_x8(t_lane + 1 * 16, az);
                    tmem_ld_x8(t_lane + 2 * 16, an);
                    if (FUSE_X) tmem_ld_x8(t_lane + 3 * 16, ax);
                } else {
                    tmem_ld_x4(t_lane + 0 * 16, ar);
                    tmem_ld_x4(t_lane + 1 * 16, az);
                    tmem_ld_x4(t_lane + 2 * 16, an);
                    if (FUSE_X) tmem_ld_x4(t_lane + 3 * 16, ax);
                }
                tmem_ld_wait();
                if (!FUSE_X) gptr += gstep;                    // block of the next time step
                if (OUT_TILES) tb = o16 + (orow >> 7) * (int64_t)(XT_TILE_BYTES / 2) + (orow & (XT_ROWS - 1)) * 8;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    const F2 accr = f2_make(__uint_as_float(ar[2 * q]), __uint_as_float(ar[2 * q + 1]));
                    const F2 accz = f2_make(__uint_as_float(az[2 * q]), __uint_as_float(az[2 * q + 1]));
                    const F2 pre_r = f2_add(FUSE_X ? br2 : g2[0][q], accr);
                    const F2 pre_z = f2_add(FUSE_X ? bz2 : g2[1][q], accz);
                    float a0, a1, b0, b1;
                    f2_get(pre_r, a0, a1);   // (pre-scaled by -log2 e, common.cuh gate_scale)
                    f2_get(pre_z, b0, b1);
                    const F2 ea = f2_make(ex2_approx(fminf(a0, EXP_CLAMP)), ex2_approx(fminf(a1, EXP_CLAMP)));
                    eb2[q] = f2_make(ex2_approx(fminf(b0, EXP_CLAMP)), ex2_approx(fminf(b1, EXP_CLAMP)));
                    r2[q] = rcp_neg_fma2(f2_fma(ea, negone2, negone2), one2);    // r = 1 / (1 + ea)
                    nzb2[q] = f2_fma(eb2[q], negone2, negone2);                   // -(1 + eb): see the tanh below
                }
            }
            REC_STAMP(8);
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const F2 accn = f2_make(__uint_as_float(an[2 * q]), __uint_as_float(an[2 * q + 1]));
                const F2 gin = FUSE_X ? f2_add(bn2, f2_make(__uint_as_float(ax[2 * q]), __uint_as_float(ax[2 * q + 1])))
                                      : g2[2][q];
                // n = tanh(x), x = gi_n + r * (gh_n + b_hn); with en = e^{2x}: n = (en - 1) / (en + 1); z = 1 / (1 + eb).
                // h = (1 - z) * n + z * h_prev = (eb * (en - 1) + h_prev * (en + 1)) / ((1 + eb) * (en + 1)):
                // ONE reciprocal for the z sigmoid and the tanh together, done on the FMA pipe (rcp_neg_fma2); the
                // exponentials are clamped at 2^60, so the denominator stays below 2^121.  Same ~2.5e-7 absolute error
                // as ATen's (h_prev - n) * z + n evaluated with approximate exp/rcp (checked against float64).
                float t0, t1;
                f2_get(f2_fma(r2[q], f2_add(accn, bhn2), gin), t0, t1);   // (pre-scaled by 2 log2 e)
                const F2 en = f2_make(ex2_approx(fminf(t0, EXP_CLAMP)), ex2_approx(fminf(t1, EXP_CLAMP)));
                const F2 enp = f2_add(en, one2);
                const F2 inv = rcp_neg_fma2(f2_mul(enp, nzb2[q]), one2);
                const F2 h2 = f2_mul(f2_fma(hprev2[q], enp, f2_mul(f2_add(en, negone2), eb2[q])), inv);
                hprev2[q] = h2;
                // fp16 hi/lo split by TRUNCATION: hi = h with the low 13 mantissa bits cleared (one LOP3; exactly an fp16
                // value for |h| >= 2^-14), lo = fp16(h - hi) from one packed subtract.  Same 22+ significant bits as the
                // round-to-nearest split (|lo| < 2^-10 |h| instead of 2^-11 |h|), 5 FMA-pipe instructions fewer per pair.
                float h0v, h1v;
                f2_get(h2, h0v, h1v);
                const float t0v = __uint_as_float(__float_as_uint(h0v) & 0xFFFFE000u);
                const float t1v = __uint_as_float(__float_as_uint(h1v) & 0xFFFFE000u);
                float l0v, l1v;
                f2_get(f2_fma(f2_make(t0v, t1v), negone2, h2), l0v, l1v);
                const __half2 hi2 = __floats2half2_rn(t0v, t1v), lo2 = __floats2half2_rn(l0v, l1v);
                hh[2 * q] = __low2half(hi2);
                hh[2 * q + 1] = __high2half(hi2);
                const __half lo0 = __low2half(lo2), lo1 = __high2half(lo2);
                hl[2 * q] = lo0;
                hl[2 * q + 1] = lo1;
                uint8_t *hw = hrow + (((int)step + 1) & 1) * L::HB;                     // (LOGITS: the buffer not being read)
                *reinterpret_cast<__half *>(hw + (2 * q) * 16) = hh[2 * q];             // B operand of the next step
                *reinterpret_cast<__half *>(hw + RT_HPLANE + (2 * q) * 16) = lo0;
                *reinterpret_cast<__half *>(hw + (2 * q + 1) * 16) = hh[2 * q + 1];
                *reinterpret_cast<__half *>(hw + RT_HPLANE + (2 * q + 1) * 16) = lo1;
            }
            if (FUSE_X && xown && more) {
                // stage x_{step+1} (loaded a step ago) into the other x buffer
                __half hi, lo;
                split_f16(xreg, hi, lo);
                uint8_t *xd = xdst + (((step + 1) & 1) ? RT_XBUF : 0);
                *reinterpret_cast<__half *>(xd) = hi;
                *reinterpret_cast<__half *>(xd + RT_XPLANE) = lo;
            }
            // publish the next B operand first: everything below (global stores of this step's output, loads of the
            // next step's pre-activations / features) is off the MMA -> gate -> MMA critical path
            REC_STAMP(9);
            fence_proxy_async_smem();     // h / x tile writes -> visible to the MMA's async-proxy reads
            tc_fence_before_sync();       // order our tcgen05.ld before the next MMA overwrites the accumulators
            if (RT_NB) {
                if (more) {   // nobody waits for the tile written by the last step: leave no arrivals pending at exit
                    if (tile == 0) named_bar_arrive<RT_BAR_H, HCOUNT>(); else named_bar_arrive<RT_BAR_H + 1, HCOUNT>();
                } else if (LOGITS) {
                    named_bar_arrive<6, 32 * RT_GATE_WARPS + 32>();   // ... except the issuer, for the last position's logits
                }
            } else {
                mbar_arrive(&h_ready[tile]);
            }
            REC_STAMP(11);
            if (FUSE_X && xok && step + 2 < T) {
                // feature value staged during the NEXT step (for the step after it); issued before this step's
                // output stores so that it is not queued behind them
                xreg = *xnext;
                xnext += xadv;
            }
            if (TRACE && trace && lane == 0 && blockIdx.x == 0 && blockIdx.y == 0 && step >= RT_TRACE_STEP0 &&
                step < RT_TRACE_STEP0 + RT_TRACE_STEPS)     // slots 16..31: when each gate warp arrived
                trace[(step - RT_TRACE_STEP0) * RT_TRACE_SLOTS + 16 + warp] = (unsigned long long)clock64();
            if (tile_ok) {
                if (!BULK_OUT && !LOGITS) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        if (OUT_TILES) {
                            tb[(2 * q) * 8] = hh[2 * q];
                            tb[XT_PLANE_BYTES / 2 + (2 * q) * 8] = hl[2 * q];
                            tb[(2 * q + 1) * 8] = hh[2 * q + 1];
                            tb[XT_PLANE_BYTES / 2 + (2 * q + 1) * 8] = hl[2 * q + 1];
                        } else {
                            float h0v, h1v;
                            f2_get(hprev2[q], h0v, h1v);
                            o32[(2 * q) * H2] = h0v;
                            o32[(2 * q + 1) * H2] = h1v;
                        }
                    }
                }
                REC_STAMP(10);   // (trace: reuses the slot of the proxy fence stamp) output stores queued
                // software pipeline: the pre-activations of the NEXT step; the loads complete under its MMAs
                if (!FUSE_X && !GI_SMEM && more) load_gi();
            }
            o32 += tstep * H2;
            orow += tstep;
            REC_STAMP(12);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (BULK_OUT && warp == RT_GATE_WARPS + 2) {
        // the tile written by the last step (its writers fenced it for the async proxy before the barrier above)
        if (elect_one()) {
            const int64_t orow = ((int64_t)blockIdx.x * T + (dir ? 0 : (T - 1))) * WT;
            uint8_t *dst = reinterpret_cast<uint8_t *>(h_out) + (orow >> 7) * (int64_t)XT_TILE_BYTES +
                           (int64_t)(dir * (H / 8)) * (XT_ROWS * 16) + (orow & (XT_ROWS - 1)) * 16;
            for (int plane = 0; plane < 2; ++plane)
                for (int kg = 0; kg < H / 8; ++kg)
                    bulk_s2g(dst + plane * XT_PLANE_BYTES + kg * (XT_ROWS * 16),
                             smem + L::h_off + plane * RT_HPLANE + kg * RT_KG, WT * 16);
            bulk_commit_group();
            bulk_wait_all();
        }
        __syncwarp();
    }
    if (warp == RT_GATE_WARPS) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

static unsigned long long *g_rec_trace = nullptr;   // device buffer [2 layers][steps][slots]; null = tracing off

unsigned long long *rec_trace_buffer() { return g_rec_trace; }

cudaError_t rec_trace_control(int enable, unsigned long long *host_out) {
    const size_t bytes = 2 * RT_TRACE_STEPS * RT_TRACE_SLOTS * sizeof(unsigned long long);
    if (host_out && g_rec_trace) {
        cudaError_t e = cudaMemcpy(host_out, g_rec_trace, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return e;
    }
    if (enable && !g_rec_trace) {
        cudaError_t e = cudaMalloc(&g_rec_trace, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemset(g_rec_trace, 0, bytes);
    }
    if (!enable && g_rec_trace) {
        cudaFree(g_rec_trace);
        g_rec_trace = nullptr;
    }
    return cudaSuccess;
}

bool rec_tc_can_fuse_logits(int64_t B, int sm_count) {
    const int64_t tiles = (B + RT_N - 1) / RT_N;
    return tiles > 0 && tiles * NDIR <= (int64_t)sm_count;      // one tile per CTA (NT == 1)
}

cudaError_t launch_rec_tc(const float *gi, const RecXArgs *fuse, const __half *w_hh_tm, const float *b_hn,
                          void *h_out, int out_tiles, int64_t B, int64_t T, int sm_count, cudaStream_t s,
                          const __half *lin_w_tc, float *plog, uint32_t prod_mask) {
    if (B == 0 || T == 0) return cudaSuccess;
    prod_mask = (prod_mask & 7u) | 1u;
    const int64_t tiles = (B + RT_N - 1) / RT_N;
    // ping-pong (2 tiles per CTA) only pays once there are more tiles than SMs to run them one per CTA
    const bool two = tiles * NDIR > (int64_t)sm_count;
    RecX xin{nullptr, nullptr, nullptr, 0};
    if (fuse) xin = RecX{fuse->feats, fuse->w_x, fuse->bias, fuse->F};
    cudaError_t e;
    unsigned long long *trace = nullptr;
#define MDK_LAUNCH_REC_T(NTV, OT, FX, TR, LG)                                                                \
    do {                                                                                                     \
        auto kern = rec_tc_kernel<NTV, OT, FX, TR, LG>;                                                      \
        constexpr int smem_bytes = RecCfg<NTV, FX, LG>::total;                                               \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);             \
        if (e != cudaSuccess) return e;                                                                      \
        dim3 grid((unsigned)((tiles + NTV - 1) / NTV), NDIR);                                                \
        kern<<<grid, RT_THREADS, smem_bytes, s>>>(gi, xin, w_hh_tm, b_hn, h_out, B, T, trace, lin_w_tc, plog, prod_mask); \
    } while (0)
#define MDK_LAUNCH_REC(NTV, OT, FX) MDK_LAUNCH_REC_T(NTV, OT, FX, false, false)
    if (lin_w_tc) {
        // layer 1 with the linear head fused in (partial logits instead of h1)
        if (fuse || out_tiles || two || !plog) return cudaErrorInvalidValue;
        if (g_rec_trace && T >= RT_TRACE_STEP0 + RT_TRACE_STEPS) {
            trace = g_rec_trace + RT_TRACE_STEPS * RT_TRACE_SLOTS;
            MDK_LAUNCH_REC_T(1, false, false, true, true);
        } else {
            MDK_LAUNCH_REC_T(1, false, false, false, true);
        }
        return cudaGetLastError();
    }
    if (g_rec_trace && !two && T >= RT_TRACE_STEP0 + RT_TRACE_STEPS) {
        // diagnostics: the two shapes the engine uses at NT = 1 (fused layer 0 -> tiles, layer 1 -> fp32 rows)
        if (fuse && out_tiles) {
            trace = g_rec_trace;
            MDK_LAUNCH_REC_T(1, true, true, true, false);
            return cudaGetLastError();
        }
        if (!fuse && !out_tiles) {
            trace = g_rec_trace + RT_TRACE_STEPS * RT_TRACE_SLOTS;
            MDK_LAUNCH_REC_T(1, false, false, true, false);
            return cudaGetLastError();
        }
    }
    if (fuse) {
        if (!out_tiles) return cudaErrorInvalidValue;   // the fused projection is layer 0, which feeds the GEMM
        if (two) MDK_LAUNCH_REC(2, true, true); else MDK_LAUNCH_REC(1, true, true);
    } else if (two) {
        if (out_tiles) MDK_LAUNCH_REC(2, true, false); else MDK_LAUNCH_REC(2, false, false);
    } else {
        if (out_tiles) MDK_LAUNCH_REC(1, true, false); else MDK_LAUNCH_REC(1, false, false);
    }
#undef MDK_LAUNCH_REC
#undef MDK_LAUNCH_REC_T
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
constexpr int GT_STAGES = 6;
constexpr int GT_KSLICE = 64;                                   // K per stage
constexpr int GT_SLICE_BYTES = XT_ROWS * GT_KSLICE * 2;         // one plane of one slice: 16 KiB
constexpr int GT_STAGE_BYTES = 2 * GT_SLICE_BYTES;              // hi + lo
constexpr int GT_BAR_OFF = GT_STAGES * GT_STAGE_BYTES;          // 192 KiB of activation stages
constexpr int GT_SMEM = GT_BAR_OFF + 256;
constexpr uint32_t GT_TMEM_COLS = 512;   // whole TMEM: the allocation then starts at column 0 (literal addresses)
constexpr uint32_t GT_W_COLS = 2 * (H2 / 2);                    // weight block hi+lo as TMEM A operand: 256 columns

// The 128x256 weight block (hi+lo) is the A operand and lives in TENSOR MEMORY (256 columns) for the CTA's whole
// life; the two 128-column accumulators take the other half.  Shared memory then only streams activations, which
// halves the tensor core's shared-memory read traffic (in SS mode an M128 N128 K16 MMA needs 8 KiB of smem reads
// per 64 cycles = the entire 128 B/clk, leaving nothing for the bulk-copy writes) and doubles the pipeline depth.
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const uint8_t *__restrict__ x_tiles, const __half *__restrict__ w_in_tm,
               const float *__restrict__ bias, float *__restrict__ gi, int64_t P, int64_t ntiles, uint32_t prod_mask,
               int *__restrict__ tile_ctr) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + GT_BAR_OFF);
    uint64_t *empty = full + GT_STAGES;
    uint64_t *acc_full = empty + GT_STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);
    // Tiles are CLAIMED, not strided: the producer takes the next tile of its weight block from tile_ctr[blk] and hands
    // the index down the pipeline (stage_tile with the stage's full barrier, acc_tile with the accumulator's).  A CTA
    // that reaches its SM late - the other lane's recurrent kernel holds up to half of them - then simply takes fewer
    // tiles, and the kernel ends when the work does instead of when the last-placed CTA has finished a fixed share.
    volatile int *stage_tile = reinterpret_cast<volatile int *>(tmem_slot + 1);   // [GT_STAGES]; -1 = no more tiles
    volatile int *acc_tile = stage_tile + GT_STAGES;                              // [2]
    volatile int *first_tile = acc_tile + 2;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    // The weight block is the FAST grid dimension: the six CTAs that stream the same activation tiles then have consecutive
    // block ids and land next to each other (same GPC / die), so five of the six reads of a tile hit L2.  With the tile index
    // fast (the round-1 layout) the six sat 24 block ids apart, spread over both dies, and DRAM read 30.3 GB per launch
    // instead of 12.5 GB (11.4 algorithmic); the kernel is at the power cap, so the saved HBM energy is clock for the whole
    // step: 10.7 -> 9.65 ms for this kernel and +7 % end to end (profiles/r01h_gemm_cta_order.md).
    const int blk = blockIdx.x;                 // weight block: dir*3 + gate

    if (tid == 0) {
        for (int i = 0; i < GT_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
        fence_mbar_init();
        first_tile[0] = atomicAdd(tile_ctr + blk, 1);      // a CTA placed after the work ran out leaves at once
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, GT_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const bool have_work = first_tile[0] < ntiles;

    if (have_work && warp >= 2) {
        // weights (row-major fp16 [blk][part][row j][k 256]) -> TMEM: lane j, plane p chunk ks at column (p*16+ks)*8
        const int q = warp & 3;
        const int jrow = q * 32 + lane;
        const uint32_t t_w = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int p = 0; p < 2; ++p) {
            const uint4 *src = reinterpret_cast<const uint4 *>(w_in_tm + (((size_t)blk * 2 + p) * H + jrow) * H2);
#pragma unroll 4
            for (int ks = 0; ks < H2 / 16; ++ks) {
                const uint4 lo4 = src[2 * ks], hi4 = src[2 * ks + 1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + (uint32_t)((p * (H2 / 16) + ks) * 8), v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    if (!have_work) {
        // nothing claimed: fall through to the TMEM release
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t tile = first_tile[0];; tile = atomicAdd(tile_ctr + blk, 1)) {
                if (tile >= ntiles) {
                    const uint32_t stage = it % GT_STAGES;
                    mbar_wait(&empty[stage], ((it / GT_STAGES) & 1) ^ 1);
                    stage_tile[stage] = -1;
                    mbar_arrive(&full[stage]);
                    break;
                }
                const uint8_t *src = x_tiles + tile * (int64_t)XT_TILE_BYTES;
                for (int s = 0; s < XT_K / GT_KSLICE; ++s, ++it) {
                    const uint32_t stage = it % GT_STAGES;
                    mbar_wait(&empty[stage], ((it / GT_STAGES) & 1) ^ 1);
                    uint8_t *dst = smem + stage * GT_STAGE_BYTES;
                    stage_tile[stage] = (int)tile;
                    mbar_arrive_expect_tx(&full[stage], GT_STAGE_BYTES);
                    bulk_g2s(dst, src + (size_t)s * GT_SLICE_BYTES, GT_SLICE_BYTES, &full[stage]);
                    bulk_g2s(dst + GT_SLICE_BYTES, src + XT_PLANE_BYTES + (size_t)s * GT_SLICE_BYTES,
                             GT_SLICE_BYTES, &full[stage]);
                }
            }
        }
    } else if (warp == 1) {
        // MMA issuer: operands kept in uniform registers (literal TMEM addresses, descriptors derived from the
        // constant dynamic-smem base, elect.sync predicate) - see the note in rec_tc_kernel.
        const uint32_t idesc = make_idesc_f16(128, XT_ROWS);
        const uint64_t b_desc0 = make_smem_desc(smem_u32(smem), XT_ROWS * 16, 128);
        if (tmem_base != 0u) {
            if (lane == 0) printf("mdk: unexpected TMEM base %u for a 512-column allocation\n", tmem_base);
            __trap();
        }
        uint32_t it = 0, tcount = 0;
        for (bool more = true; more; ++tcount) {
            const uint32_t as = tcount & 1;
            mbar_wait(&acc_empty[as], ((tcount >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            for (int s = 0; s < XT_K / GT_KSLICE; ++s, ++it) {
                const uint32_t stage = it % GT_STAGES;
                mbar_wait(&full[stage], (it / GT_STAGES) & 1);
                tc_fence_after_sync();
                if (s == 0) {
                    const int tile = stage_tile[stage];
                    if (lane == 0) {
                        acc_tile[as] = tile;
                        fence_proxy_async_smem();      // the index is read behind a barrier the tensor core arrives on
                        if (tile < 0) mbar_arrive(&acc_full[as]);
                    }
                    __syncwarp();
                    if (tile < 0) { more = false; break; }
                }
                if (elect_one()) {
                    const uint32_t d = GT_W_COLS + as * XT_ROWS;
                    const uint64_t b_s = b_desc0 + (uint64_t)((stage * GT_STAGE_BYTES) >> 4);
#pragma unroll
                    for (int prod = 0; prod < 3; ++prod) {
                        if (prod && !(prod_mask & (1u << prod))) continue;
                        const int pa = (prod == 2) ? 1 : 0;   // W part
                        const int pb = (prod == 1) ? 1 : 0;   // x part
#pragma unroll
                        for (int ks = 0; ks < GT_KSLICE / 16; ++ks) {
                            const uint32_t a_t = (uint32_t)((pa * (H2 / 16) + s * (GT_KSLICE / 16) + ks) * 8);
                            const uint64_t bd = b_s + (uint64_t)((pb * GT_SLICE_BYTES + ks * 2 * (XT_ROWS * 16)) >> 4);
                            umma_f16_ts(d, a_t, bd, idesc, (s | prod | ks) ? 1u : 0u);
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
        for (uint32_t tcount = 0;; ++tcount) {
            const uint32_t as = tcount & 1;
            mbar_wait(&acc_full[as], (tcount >> 1) & 1);
            const int64_t tile = acc_tile[as];
            if (tile < 0) break;
            tc_fence_after_sync();
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + GT_W_COLS + as * XT_ROWS;
            // quad layout (common.cuh): column c of this tile = tile-step tile*8 + c/16, window c%16; the four windows
            // of a quad are one 16-byte store, contiguous over the warp's 32 rows j
            float4 *out = reinterpret_cast<float4 *>(gi) + (tile * (XT_ROWS / WT)) * (int64_t)(GI_TS_FLOATS / 4) +
                          (int64_t)(blk * 4) * H + j;
            const int64_t prem = P - tile * XT_ROWS;   // rows of this tile that exist (a multiple of 16)
#pragma unroll 1
            for (int c32 = 0; c32 < XT_ROWS; c32 += 32) {
                uint32_t v[32];
                tmem_ld_x32(t_lane + c32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const int c = c32 + i;
                    if (c < prem)
                        st_stream4(out + (int64_t)(c >> 4) * (GI_TS_FLOATS / 4) + ((c & 15) >> 2) * H,
                                   make_float4(__uint_as_float(v[i]) + bj, __uint_as_float(v[i + 1]) + bj,
                                               __uint_as_float(v[i + 2]) + bj, __uint_as_float(v[i + 3]) + bj));
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

cudaError_t launch_gemm_tc(const void *x_tiles, const __half *w_in_tm, const float *bias, float *gi, int64_t P,
                           int sm_count, cudaStream_t s, uint32_t prod_mask, int *tile_ctr) {
    if (P == 0) return cudaSuccess;
    const int64_t ntiles = (P + XT_ROWS - 1) / XT_ROWS;
    // (the attribute is per device: set it on every launch, a process may drive several GPUs)
    cudaError_t ea = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM);
    if (ea != cudaSuccess) return ea;
    int64_t ct = sm_count / 6;
    static const int rows_env = getenv("MDK_GEMM_ROWS") ? atoi(getenv("MDK_GEMM_ROWS")) : 0;   // experiments only
    if (rows_env > 0 && rows_env < ct) ct = rows_env;
    if (ct < 1) ct = 1;
    if (ct > ntiles) ct = ntiles;
    dim3 grid(6, (unsigned)ct);
    ea = cudaMemsetAsync(tile_ctr, 0, 6 * sizeof(int), s);     // the six per-weight-block tile counters
    if (ea != cudaSuccess) return ea;
    gemm_tc_kernel<<<grid, GT_THREADS, GT_SMEM, s>>>(reinterpret_cast<const uint8_t *>(x_tiles), w_in_tm, bias, gi,
                                                     P, ntiles, (prod_mask & 7u) | 1u, tile_ctr);
    return cudaGetLastError();
}

// =====================================================================================================
// Self test of the UMMA building block: D[128][N] = A[128][K] . B[N][K]^T, fp16 hi/lo split, one CTA.
// variant 0 = SS-mode descriptors; 1 = LBO/SBO swapped; 2 = descriptor version bits cleared;
// 3 = A operand from TMEM (production layout of the recurrent kernel); 4 = same with the fp16 pair order swapped.
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
    if (warp == 0) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const bool a_tmem = variant >= 3;
    if (a_tmem) {
        // A (both planes) -> TMEM columns [128, 128 + K): plane p, chunk ks at column 128 + (p*(K/16) + ks)*8
        const int r = warp * 32 + lane;
        for (int p = 0; p < 2; ++p)
            for (int ks = 0; ks < K / 16; ++ks) {
                uint32_t v[8];
                for (int c = 0; c < 8; ++c) {
                    const int k0 = ks * 16 + 2 * c;
                    const uint32_t e0 = *reinterpret_cast<const uint16_t *>(sa + p * a_plane + (k0 / 8) * (128 * 16) + r * 16 + (k0 % 8) * 2);
                    const uint32_t e1 = *reinterpret_cast<const uint16_t *>(sa + p * a_plane + ((k0 + 1) / 8) * (128 * 16) + r * 16 + ((k0 + 1) % 8) * 2);
                    v[c] = (variant == 4) ? ((e0 << 16) | e1) : ((e1 << 16) | e0);
                }
                tmem_st_x8(tmem_base + ((uint32_t)(warp * 32) << 16) + 128 + (uint32_t)((p * (K / 16) + ks) * 8), v);
            }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        uint32_t acc = 0;
        for (int prod = 0; prod < 3; ++prod) {
            const int pa = (prod == 2), pb = (prod == 1);
            for (int ks = 0; ks < K / 16; ++ks) {
                if (a_tmem) {
                    const uint64_t bd = make_smem_desc(smem_u32(sb + pb * b_plane) + ks * 2 * N * 16, N * 16, 128);
                    umma_f16_ts(tmem_base, tmem_base + 128 + (uint32_t)((pa * (K / 16) + ks) * 8), bd, idesc, acc);
                    acc = 1;
                    continue;
                }
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
    if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
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
