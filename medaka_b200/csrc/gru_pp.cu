// Two-tile ("ping-pong") GRU recurrence on tcgen05 for sm_100a.
//
// Reference arithmetic: torch.nn.GRU as medaka/architectures/gru.py:46-52,66 uses it (gate order r, z, n;
// n = tanh(gi_n + r * (gh_n + b_hn)); h = (1 - z) * n + z * h); the linear head of gru.py:53-55,67 rides along in layer 1.
//
// Why a second kernel.  rec_tc_kernel (gru_tc.cu) runs ONE tile of 16 windows per CTA: a time step is a single dependent
// chain  h published -> 72 MMAs -> gate math -> h published  (~1 900 cycles, of which the tensor pipe works ~700 and the
// issue slots ~700: profiles/r01i_rec_cycle_trace.txt).  Nothing can overlap because there is nothing else to do.  This
// kernel gives a CTA TWO independent tiles (A, B) of the same direction and lets the chains interleave: while the gate
// warps of A do their sigmoid / tanh work the tensor pipe runs B's MMAs and vice versa.
//
//   warps 0-7    gate warps of tile A   (warp w: TMEM lane quarter w % 4, windows (w / 4) * 8 .. + 7: 8 windows / thread)
//   warps 8-15   gate warps of tile B
//   warp 17, 18  MMA issuers of tile A / B for the r and n blocks (one elected thread each)
//   warp 19, 20  MMA issuers of tile A / B for the z block and (layer 1) the logits.  Two issuers per tile because issue is
//                what limits the MMA phase here: an MMA costs ~7 issue slots of descriptor arithmetic (ptxas routes the
//                operands through vector registers and R2UR), and unlike in the one-tile kernel the issuer shares its
//                scheduler with gate warps of the OTHER tile that are busy - measured 21 cycles per MMA from one issuer
//                against 9.75 on the tensor pipe (profiles/r02_pp_bringup.md).  Blocks are never split between warps, so
//                "first MMA of a block overwrites the accumulator" stays well defined.
//   warp 21, 22  relay of tile A / B: the only waiter on that tile's commit mbarriers; releases the gate warps through
//                named hardware barriers (a blocked mbarrier.try_wait of 8 warps costs 100-250 cycles more than one
//                bar.sync, measured in round 1); in its idle time it stages the next pre-activation block (layer 1: one
//                24 KiB bulk copy into shared memory, two steps ahead) and prefetches L2
//   warp 16      (layer 1) reads the 5 x 16 partial logits out of tensor memory and writes them
//   warp 16, 21  (layer 0) copy warp of tile A / B: moves the published h tile - which already is the operand-tile image
//                the projection GEMM wants - to HBM with 16-byte loads and stores.  (Round 2, first cut: 32 bulk copies
//                of 256 B per tile-step from the relay; the copy engine became the limiter of the whole kernel at
//                ~70 cycles per small copy, 4 800 cycles per pair of tile-steps: profiles/r02_pp_bringup.md)
//
// Tensor memory (512 columns, all of it):  W_hh fp16 hi|lo as the A operand [0, 384);
//   layer 0:  W_ih (K = 16) hi|lo [384, 432);  accumulators r, z, n, W_in.x [432, 496)
//   layer 1:  accumulators r, z, n [384, 432);  logits [432, 448);  W_lin hi plane [448, 512)
// There is ONE set of accumulators and the two tiles take turns: the tensor pipe executes MMAs in issue order, tile B's
// r block can only start after tile A's n block, and by then A's gate warps have long since read A's r block.  That
// "long since" is made a guarantee by the `free_acc[X][g]` mbarriers ("block g may be used by tile X"): the gate warps
// of the OTHER tile arrive on it after their tcgen05.ld of block g, and issuer X waits for that arrival before it queues
// its own block g.  Uses of a block strictly alternate A, B, A, B (each tile's next use needs the other's consumption);
// every barrier has one waiter that asks for consecutive phases, so the parity wait is unambiguous.  Sharing is what makes everything fit: with per-tile accumulators neither W_ih (layer 0) nor W_lin
// (layer 1) could stay resident, and an SS-mode MMA (A operand from shared memory) costs 40 cycles instead of 10.
//
// Per-step protocol of tile X (named barriers H_X, R_X, Z_X, N_X; all counts = 8 gate warps + 1 service warp):
//   gate warps: ... write h_t (fp16 hi/lo, K-major B-operand image, double buffered) -> fence.proxy.async -> arrive H_X
//   issuer X:   sync H_X -> [wait free_acc[X][r]] x-part of n (layer 0), r block, commit -> [free_acc[X][z]] z block,
//               commit -> [free_acc[X][n]] n block, commit -> (layer 1) [log_free[X]] logits of h_{t-1}, commit
//   relay X:    wait commit r (+ gi landed) -> arrive R_X;  wait commit z -> arrive Z_X;  wait commit n -> arrive N_X
//   gate warps: sync R_X -> tcgen05.ld r (+x) -> arrive free_acc[other][r] -> sigmoid(r), r * b_hn + gi_n
//               sync Z_X -> ld z -> arrive free_acc[other][z] -> e^-z and the z-dependent halves of the h update
//               sync N_X -> ld n -> arrive free_acc[other][n] -> tanh, h_t, split, store
// Every mbarrier wait is bounded and traps (ptx.cuh); the named barriers cannot be bounded, so bring-up runs go under
// `timeout`.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"
#include "rec_common.cuh"

namespace mdk {

constexpr int PP_TILE_WARPS = 8;
constexpr int PP_GATE_WARPS = 2 * PP_TILE_WARPS;
// service warps.  Warp 16 must be the one that reads the logits (tcgen05.ld reaches TMEM lanes 32 * (warp % 4) ..., the
// classes sit in lanes 0..4); the two issuers - the only service warps with real instruction streams - get a scheduler
// each that they share with nobody but gate warps (scheduler = warp % 4)
constexpr int PP_W_AUX0 = 16;     // layer 1: logits warp; tiles out: copy warp of tile A
constexpr int PP_W_ISS = 17;      // 17, 18: issuers (r and n blocks) of tile A / B; 19, 20: issuers (z block, logits) of A / B
constexpr int PP_W_REL = 21;      // 21, 22: relay of tile A / B
constexpr int PP_W_AUX1 = 23;     // tiles out: copy warp of tile B
constexpr int PP_WARPS = 24;
constexpr int PP_THREADS = 32 * PP_WARPS;                 // 768 -> at most 80 registers per thread
constexpr int PP_BAR_H = 1, PP_BAR_R = 3, PP_BAR_Z = 5, PP_BAR_N = 7, PP_BAR_FIN = 9;   // + X
constexpr int PP_NB = 32 * (PP_TILE_WARPS + 1);           // 8 gate warps + 1 service warp
constexpr int PP_NB2 = 32 * (PP_TILE_WARPS + 2);          // 8 gate warps + 2 service warps
constexpr int PP_NB3 = 32 * (PP_TILE_WARPS + 3);          // 8 gate warps + 3 service warps
constexpr int PP_GI_BUFS = 3;
constexpr int PP_GI_BLOCK = (GI_TS_FLOATS / 2) * 4;       // one (tile-step, direction) of gi: 24 576 bytes
constexpr int PP_WL_PLANE = 16 * 64 * 16;                 // W_lin lo plane as an M = 64 shared-memory A operand: 16 KiB

template <int LAYER>
struct PPCfg {
    // LAYER 0: layer 0 with the input projection fused (features in, operand tiles out)
    // LAYER 1: layer 1 (gi in, partial logits out)
    // LAYER 2: layer 0 of a model with F > 16 features (gi from inproj0_kernel in, operand tiles out)
    static constexpr bool IN_X = LAYER == 0;
    static constexpr bool OUT_LOG = LAYER == 1;
    static constexpr uint32_t wx_col = RT_WT_COLS;                              // IN_X
    static constexpr uint32_t acc_col = RT_WT_COLS + (IN_X ? RT_WX_COLS : 0);   // r, z, n (, x): 16 columns each
    static constexpr uint32_t log_col = acc_col + 48;                           // OUT_LOG
    static constexpr uint32_t wl_col = 448;                                     // OUT_LOG: W_lin hi plane, 64 columns
    static_assert(IN_X ? acc_col + 64 <= 512 : log_col + 16 <= wl_col, "TMEM budget");
    // shared memory: h tiles [tile 2][buffer 2][plane 2][RT_HPLANE]; layer 0: x tiles [tile 2][buffer 2][RT_XBUF];
    // layer 1: gi staging [tile 2][PP_GI_BUFS][PP_GI_BLOCK], W_lin lo plane
    static constexpr int h_off = 0;
    static constexpr int h_end = 2 * 2 * 2 * RT_HPLANE;
    static constexpr int x_off = h_end;
    static constexpr int gi_off = h_end;
    static constexpr int wl_off = gi_off + 2 * PP_GI_BUFS * PP_GI_BLOCK;
    static constexpr int data_end = IN_X ? x_off + 2 * 2 * RT_XBUF : wl_off + (OUT_LOG ? PP_WL_PLANE : 0);
    static constexpr int bar_off = ((data_end + 127) / 128) * 128;
    // mbarriers: acc[2][3], gi_full[2][3], free_acc[2][3], log_full, log_free[2] ; then the TMEM slot
    static constexpr int n_bars = 6 + 6 + 6 + 1 + 2;
    static constexpr int tmem_off = bar_off + n_bars * 8;
    static constexpr int end_ = tmem_off + 16;
    // the CTA owns all 512 TMEM columns of its SM: ask for more than half of the shared memory so that a second CTA can
    // never become co-resident (it would spin in tcgen05.alloc)
    static constexpr int total = end_ > 120 * 1024 ? end_ : 120 * 1024;
    static_assert(total <= 227 * 1024, "smem budget");
};

struct PPArgs {
    const float *gi;          // layer 1: pre-activations, quad layout (common.cuh)
    RecX xin;                 // layer 0: fused input projection
    const __half *w_hh;       // [dir][hi/lo][gate][row 128][k 128]
    const float *b_hn;        // [dir][128], pre-scaled
    void *h_out;              // layer 0: fp16 hi/lo operand tiles of the projection GEMM
    int64_t B, T, ntiles;
    unsigned long long *trace;
    const __half *lin_w_tc;   // layer 1: packed W_lin (misc.cu pack_linear_kernel)
    float *plog;              // layer 1: partial logits [dir][tile][t][class 5][16 windows]
    uint32_t prod_mask;       // bit 0: W_hi.h_hi, bit 1: W_hi.h_lo, bit 2: W_lo.h_hi (7 = fp32-faithful)
    uint32_t debug;           // TRACE instantiations only: switch parts of the step off to see what they cost
                              // (results are then wrong).  1: no h-tile stores, 2: no proxy fence, 4: no gate arithmetic,
                              // 8: no x staging, 16: no gi staging / prefetch, 32: no tile copy-out
};
#define PP_DBG(bit) (TRACE && (a.debug & (bit)))

constexpr int PP_TRACE_STEP0 = 512, PP_TRACE_STEPS = 16, PP_TRACE_SLOTS = 40;   // == the rec_tc trace geometry
#define PP_STAMP(slot)                                                    \
    do {                                                                  \
        if (TRACE && tr) tr[slot] = (unsigned long long)clock64();        \
    } while (0)

template <bool TRACE>
__device__ __forceinline__ unsigned long long *pp_trace_row(const PPArgs &a, int64_t step, bool leader) {
    if (!TRACE || !a.trace || !leader || blockIdx.x != 0 || blockIdx.y != 0) return nullptr;
    if (step < PP_TRACE_STEP0 || step >= PP_TRACE_STEP0 + PP_TRACE_STEPS) return nullptr;
    return a.trace + (step - PP_TRACE_STEP0) * PP_TRACE_SLOTS;
}

// ------------------------------------------------------------------------------------------------ issuer of tile X
template <int LAYER, bool ALLP, bool TRACE>
__device__ __noinline__ void pp_issuer(uint8_t *smem, const PPArgs &a, const int X, const int role) {
    using L = PPCfg<LAYER>;
    constexpr bool IN_X = L::IN_X, OUT_LOG = L::OUT_LOG;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint64_t *acc = bars + X * 3;
    uint64_t *free_acc = bars + 12 + X * 3;        // completed by the other tile's gate warps
    uint64_t *log_full = bars + 18, *log_free = bars + 19 + X;
    // Operands sit in uniform registers: literal TMEM addresses, descriptors derived from the constant dynamic-smem
    // base, issue predicated by elect.sync (otherwise every tcgen05.mma is wrapped in an R2UR waterfall loop).
    const uint32_t idesc = make_idesc_f16(128, RT_N);
    const uint32_t idesc64 = make_idesc_f16(64, RT_N);
    // descriptors as (low word, high word): only the address field of the low word moves (ptx.cuh umma_f16_ts2)
    const uint32_t d_hi = smem_desc_hi(128);
    const uint32_t h_lo0 = smem_desc_lo(smem_u32(smem + L::h_off), RT_KG) + (uint32_t)X * ((4 * RT_HPLANE) >> 4);
    const uint32_t x_lo0 = smem_desc_lo(smem_u32(smem + L::x_off), RT_KG) + (uint32_t)X * ((2 * RT_XBUF) >> 4);
    const uint32_t wl_lo0 = smem_desc_lo(smem_u32(smem + L::wl_off), 64 * 16);
    const uint32_t mask = ALLP ? 7u : a.prod_mask;
    const int64_t T = a.T;

    // W_hh[gate] . h (tile buffer hd), the selected fp16 products, into accumulator columns d; returns the accumulate flag
    auto issue_h = [&](uint32_t d, int gate, uint32_t hd, uint32_t accf) -> uint32_t {
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
            if (!ALLP && !(mask & (1u << prod))) continue;
            const int pa = (prod == 2) ? 1 : 0;   // W part: hi, hi, lo
            const int pb = (prod == 1) ? 1 : 0;   // h part: hi, lo, hi
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                umma_f16_ts2(d, (uint32_t)(((pa * 3 + gate) * 8 + ks) * 8),
                             hd + (uint32_t)((pb * RT_HPLANE + ks * 2 * RT_KG) >> 4), d_hi, idesc, accf);
                accf = 1u;
            }
        }
        return accf;
    };
    // + W_ih[gate] . x_t (layer 0)
    auto issue_x = [&](uint32_t d, int gate, uint32_t xd, uint32_t accf) -> uint32_t {
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
            if (!ALLP && !(mask & (1u << prod))) continue;
            const int pa = (prod == 2) ? 1 : 0;
            const int pb = (prod == 1) ? 1 : 0;
            umma_f16_ts2(d, L::wx_col + (uint32_t)((pa * 3 + gate) * 8), xd + (uint32_t)((pb * RT_XPLANE) >> 4), d_hi, idesc, accf);
            accf = 1u;
        }
        return accf;
    };
    // partial logits of the h in tile buffer hd: W_lin hi plane from tensor memory (M = 128, rows >= 5 zero), lo plane
    // from shared memory (M = 64: rows 0..4 land in the same TMEM lanes)
    auto issue_logits = [&](uint32_t hd) {
        uint32_t accf = 0u;
#pragma unroll
        for (int prod = 0; prod < 3; ++prod) {
            if (!ALLP && !(mask & (1u << prod))) continue;
            const int pb = (prod == 1) ? 1 : 0;
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                const uint32_t bd = hd + (uint32_t)((pb * RT_HPLANE + ks * 2 * RT_KG) >> 4);
                if (prod != 2) umma_f16_ts2(L::log_col, L::wl_col + (uint32_t)(ks * 8), bd, d_hi, idesc, accf);
                else umma_f16_ss2(L::log_col, wl_lo0 + (uint32_t)((ks * 2 * (64 * 16)) >> 4), d_hi, bd, d_hi, idesc64, accf);
                accf = 1u;
            }
        }
        umma_commit(log_full);
    };

    // role 0: the r block (with the n gate's W_in.x, layer 0) and the n block; role 1: the z block and the logits
    constexpr int NB_H = OUT_LOG ? PP_NB2 : PP_NB3;      // H_X: 8 gate warps + 2 issuers (+ the copy warp when tiles go out)
#pragma unroll 1
    for (int64_t s = 0; s < T; ++s) {
        const uint32_t par = (uint32_t)(s & 1);
        unsigned long long *tr = pp_trace_row<TRACE>(a, s, true);
        named_bar_sync_id<NB_H>(PP_BAR_H + X);
        tc_fence_after_sync();
        if (elect_one()) {
            const uint32_t hd = h_lo0 + par * ((2 * RT_HPLANE) >> 4);
            const uint32_t xd = x_lo0 + par * (RT_XBUF >> 4);
            // the use just before this one was the other tile's: B's of step s-1 for A, A's of step s for B
            const bool guard = (X == 1) || s > 0;          // the very first use of the accumulators needs no hand-over
            const uint32_t gpar = (uint32_t)((X == 0 ? s - 1 : s) & 1);
            if (role == 0) {
                PP_STAMP(X * 20 + 0);
                if (guard) mbar_wait(&free_acc[0], gpar);
                PP_STAMP(X * 20 + 15);
                tc_fence_after_sync();
                if (IN_X) issue_x(L::acc_col + 48, 2, xd, 0u);   // W_in . x of the n gate keeps its own columns
                {
                    const uint32_t f = issue_h(L::acc_col + 0, 0, hd, 0u);
                    if (IN_X) issue_x(L::acc_col + 0, 0, xd, f);
                }
                umma_commit(&acc[0]);
                PP_STAMP(X * 20 + 1);
                if (guard) mbar_wait(&free_acc[2], gpar);
                PP_STAMP(X * 20 + 17);
                tc_fence_after_sync();
                issue_h(L::acc_col + 32, 2, hd, 0u);
                umma_commit(&acc[2]);
                PP_STAMP(X * 20 + 3);
            } else {
                if (guard) mbar_wait(&free_acc[1], gpar);
                PP_STAMP(X * 20 + 16);
                tc_fence_after_sync();
                {
                    const uint32_t f = issue_h(L::acc_col + 16, 1, hd, 0u);
                    if (IN_X) issue_x(L::acc_col + 16, 1, xd, f);
                }
                umma_commit(&acc[1]);
                PP_STAMP(X * 20 + 2);
                if (OUT_LOG && s > 0) {
                    // the tile buffer holds h of the previous step: its logits ride in the shadow of the gate phase
                    // this tile's logits use k = s - 1 follows B's use k-1 (tile A) / A's use k (tile B)
                    if (X == 1 || s > 1) mbar_wait(log_free, (uint32_t)((X == 0 ? s - 2 : s - 1) & 1));
                    PP_STAMP(X * 20 + 18);
                    tc_fence_after_sync();
                    issue_logits(hd);
                    PP_STAMP(X * 20 + 14);
                }
            }
        }
        __syncwarp();
    }
    if (OUT_LOG && role == 1) {
        // h of the last step: every gate warp of the tile has published it (FIN_X), one more round of logits MMAs
        named_bar_sync_id<PP_NB>(PP_BAR_FIN + X);
        tc_fence_after_sync();
        if (elect_one()) {
            const uint32_t hd = h_lo0 + (uint32_t)(T & 1) * ((2 * RT_HPLANE) >> 4);
            if (X == 1 || T > 1) mbar_wait(log_free, (uint32_t)((X == 0 ? T - 2 : T - 1) & 1));   // use k = T - 1
            tc_fence_after_sync();
            issue_logits(hd);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ relay of tile X
template <int LAYER, bool TRACE>
__device__ __noinline__ void pp_relay(uint8_t *smem, const PPArgs &a, const int X, int lane) {
    using L = PPCfg<LAYER>;
    constexpr bool IN_X = L::IN_X, OUT_LOG = L::OUT_LOG;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint64_t *acc = bars + X * 3;
    uint64_t *gi_full = bars + 6 + X * 3;
    const int dir = blockIdx.y;
    const int64_t T = a.T;
    const int64_t tile = (int64_t)blockIdx.x * 2 + X;
    const bool tile_ok = tile < a.ntiles;
    const bool lead = lane == 0;            // one fixed thread issues and tracks all bulk copies of this warp
    uint8_t *gi_buf = smem + L::gi_off + X * PP_GI_BUFS * PP_GI_BLOCK;

    auto stage_gi = [&](int64_t sp) {       // gi block of step sp -> buffer sp % 3
        const int64_t t = dir ? (T - 1 - sp) : sp;
        const float *src = a.gi + (tile * T + t) * GI_TS_FLOATS + (int64_t)dir * (GI_TS_FLOATS / 2);
        const int b = (int)(sp % PP_GI_BUFS);
        mbar_arrive_expect_tx(&gi_full[b], PP_GI_BLOCK);
        bulk_g2s(gi_buf + b * PP_GI_BLOCK, src, PP_GI_BLOCK, &gi_full[b]);
    };
    auto prefetch_gi = [&](int64_t sp) {    // 2 KiB pieces: one 24 KiB prefetch measured as if never issued (round 1)
        const int64_t t = dir ? (T - 1 - sp) : sp;
        const float *blk = a.gi + (tile * T + t) * GI_TS_FLOATS + (int64_t)dir * (GI_TS_FLOATS / 2);
#pragma unroll
        for (int i = 0; i < 12; ++i) bulk_prefetch_l2(blk + i * 512, 2048);
    };
    auto prefetch_x = [&](int64_t step) {   // feature rows, X_PREFETCH_EVERY steps at a time, X_PREFETCH_AHEAD ahead
        if ((step & (X_PREFETCH_EVERY - 1)) != 0) return;
        const int64_t s0 = step + X_PREFETCH_AHEAD;
        const int64_t s1 = s0 + X_PREFETCH_EVERY <= T ? s0 + X_PREFETCH_EVERY : T;
        if (s0 >= T) return;
        const int64_t t_lo = dir ? (T - s1) : s0;
        const int64_t nbytes = (s1 - s0) * a.xin.F * 4;
        for (int w = 0; w < WT; ++w) {
            if (tile * WT + w >= a.B) break;
            const uintptr_t p = reinterpret_cast<uintptr_t>(a.xin.feats + ((tile * WT + w) * T + t_lo) * a.xin.F);
            const uintptr_t p0 = p & ~(uintptr_t)15;
            bulk_prefetch_l2(reinterpret_cast<const void *>(p0), (uint32_t)(((p + nbytes - p0) + 15) & ~(uintptr_t)15));
        }
    };
    const bool no_gi = PP_DBG(16);
    if (!IN_X && lead && tile_ok && !no_gi) {
        stage_gi(0);
        if (T > 1) stage_gi(1);
    }
#pragma unroll 1
    for (int64_t s = 0; s < T; ++s) {
        const uint32_t par = (uint32_t)(s & 1);
        unsigned long long *tr = pp_trace_row<TRACE>(a, s, lead);
        if (lead) {
            mbar_wait(&acc[0], par);
            if (!IN_X && tile_ok && !no_gi) mbar_wait(&gi_full[s % PP_GI_BUFS], (uint32_t)((s / PP_GI_BUFS) & 1));
        }
        __syncwarp();
        tc_fence_after_sync();
        tc_fence_before_sync();
        named_bar_arrive_id<PP_NB>(PP_BAR_R + X);
        PP_STAMP(X * 20 + 4);
        if (lead) mbar_wait(&acc[1], par);
        __syncwarp();
        tc_fence_after_sync();
        tc_fence_before_sync();
        named_bar_arrive_id<PP_NB>(PP_BAR_Z + X);
        PP_STAMP(X * 20 + 5);
        if (lead) mbar_wait(&acc[2], par);
        __syncwarp();
        tc_fence_after_sync();
        tc_fence_before_sync();
        if (OUT_LOG) named_bar_arrive_id<PP_NB>(PP_BAR_N + X);
        else named_bar_arrive_id<PP_NB2>(PP_BAR_N + X);
        PP_STAMP(X * 20 + 6);
        // ---- idle until the next commit: staging / copy-out / prefetch (one thread) ----
        if (lead && tile_ok && !no_gi) {
            if (!IN_X) {
                // buffer (s+2) % 3 == (s-1) % 3 was last read in step s-1, which every gate warp had left when it
                // arrived on H_X(s), and the MMAs of step s (just committed) were issued after that barrier
                if (s + 2 < T) stage_gi(s + 2);
                if (s + 2 + GI_PREFETCH_STEPS < T) prefetch_gi(s + 2 + GI_PREFETCH_STEPS);
            }
            if (IN_X) prefetch_x(s);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------ copy warp of tile X
// Tiles out (layer 0): the h tile the gate warps publish is the K-major operand image the projection GEMM reads
// ([plane][k-group 16][16 windows][8 halfs]); per (plane, k-group) its 256 bytes are contiguous in the GEMM's operand
// tile in HBM.  One warp moves the 8 KiB with 16-byte accesses: two 256-byte chunks per warp-wide instruction.
// It takes part in H_X (it may read the buffer the gate warps have just published) and in N_X of the same step (it
// arrives once its loads are in registers: the buffer it read is only overwritten after N_X of the NEXT step, which
// needs this warp's next arrival).
template <int LAYER, bool TRACE>
__device__ __noinline__ void pp_copy(uint8_t *smem, const PPArgs &a, const int X, int lane) {
    using L = PPCfg<LAYER>;
    const int dir = blockIdx.y;
    const int64_t T = a.T;
    const int64_t tile = (int64_t)blockIdx.x * 2 + X;
    const bool tile_ok = tile < a.ntiles && !PP_DBG(32);
    const int half = lane >> 4, l16 = lane & 15;
    auto copy = [&](int64_t sidx, int buf, bool arrive) {
        const int64_t orow = (tile * T + (dir ? (T - 1 - sidx) : sidx)) * WT;
        uint8_t *dst = reinterpret_cast<uint8_t *>(a.h_out) + (orow >> 7) * (int64_t)XT_TILE_BYTES +
                       (int64_t)(dir * (H / 8)) * (XT_ROWS * 16) + (orow & (XT_ROWS - 1)) * 16 + l16 * 16;
        const uint8_t *src = smem + L::h_off + (X * 2 + buf) * 2 * RT_HPLANE + l16 * 16;
        // two batches of eight 16-byte accesses per lane (32 registers in flight); chunk c: plane = c / 16, k-group = c % 16
#pragma unroll
        for (int b8 = 0; b8 < 2; ++b8) {
            uint4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int c = (b8 * 8 + i) * 2 + half;
                v[i] = *reinterpret_cast<const uint4 *>(src + (c >> 4) * RT_HPLANE + (c & 15) * RT_KG);
            }
            if (b8 == 1 && arrive) named_bar_arrive_id<PP_NB2>(PP_BAR_N + X);     // every load has been issued ...
            if (tile_ok) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {                                      // ... and is complete before its store
                    const int c = (b8 * 8 + i) * 2 + half;
                    *reinterpret_cast<uint4 *>(dst + (c >> 4) * XT_PLANE_BYTES + (c & 15) * (XT_ROWS * 16)) = v[i];
                }
            }
        }
    };
#pragma unroll 1
    for (int64_t s = 0; s < T; ++s) {
        named_bar_sync_id<PP_NB3>(PP_BAR_H + X);
        if (s > 0) copy(s - 1, (int)(s & 1), true);
        else named_bar_arrive_id<PP_NB2>(PP_BAR_N + X);
    }
    named_bar_sync_id<PP_NB>(PP_BAR_FIN + X);        // h of the last step
    copy(T - 1, (int)(T & 1), false);
}

// ------------------------------------------------------------------------------------------------ logits warp (layer 1)
template <bool TRACE>
__device__ __noinline__ void pp_logits(uint8_t *smem, const PPArgs &a, int lane) {
    using L = PPCfg<1>;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint64_t *log_full = bars + 18, *log_free = bars + 19;
    const int dir = blockIdx.y;
    const int64_t T = a.T;
#pragma unroll 1
    for (int64_t v = 0; v < 2 * T; ++v) {
        const int X = (int)(v & 1);
        const int64_t sidx = v >> 1;
        mbar_wait_warp(log_full, (uint32_t)(v & 1));
        tc_fence_after_sync();
        uint32_t r[16];
        tmem_ld_x16(L::log_col, r);           // this warp (16 % 4 == 0) reads TMEM lanes 0..31; lanes 0..4 = classes
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&log_free[1 - X]);     // the accumulator is free for the other tile's next logits
        const int64_t tile = (int64_t)blockIdx.x * 2 + X;
        if (tile < a.ntiles && lane < NCLS) {
            const int64_t t = dir ? (T - 1 - sidx) : sidx;
            float4 *dst = reinterpret_cast<float4 *>(a.plog + (((int64_t)dir * a.ntiles + tile) * T + t) * PLOG_TS_FLOATS +
                                                     lane * WT);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                     __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
        }
    }
}

// ------------------------------------------------------------------------------------------------ gate warps of tile X
template <int LAYER, bool TRACE>
__device__ __noinline__ void pp_gate(uint8_t *smem, const PPArgs &a, const int X, int w8, int lane) {
    using L = PPCfg<LAYER>;
    constexpr bool L0 = L::IN_X;                            // fused input projection: biases + x staging here, no gi
    constexpr int NP = 4;                                   // 8 windows per thread as 4 packed fp32 pairs
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint64_t *freed = bars + 12 + (1 - X) * 3;      // free_acc of the OTHER tile: it may use a block once we have read it
    const int dir = blockIdx.y;
    const int64_t T = a.T;
    const int q = w8 & 3, ch = w8 >> 2;
    const int j = q * 32 + lane;                            // hidden unit == TMEM lane
    const int c0 = ch * 8;                                  // first of this thread's 8 windows
    const int64_t tile = (int64_t)blockIdx.x * 2 + X;
    const bool tile_ok = tile < a.ntiles;                   // an odd tile count leaves the last CTA's tile B empty: it
                                                            // runs the protocol on zeros and touches no global memory
    const uint32_t t_lane = ((uint32_t)(q * 32) << 16) + L::acc_col + (uint32_t)c0;
    uint8_t *hrow = smem + L::h_off + X * 4 * RT_HPLANE + (j >> 3) * RT_KG + (j & 7) * 2 + c0 * 16;
    const bool elected = lane == 0;

    const F2 one2 = f2_make(1.0f, 1.0f), negone2 = f2_make(-1.0f, -1.0f);
    const float bhn = a.b_hn[dir * H + j];
    const F2 bhn2 = f2_make(bhn, bhn);
    F2 br2 = f2_make(0.f, 0.f), bz2 = br2, bn2 = br2;
    // layer 0: this thread's share of the x_t staging (16 windows x F values per tile-step)
    const int tix = w8 * 32 + lane;
    const int xF = L0 ? a.xin.F : 1;
    const int xn = tix / xF, xf = tix - xn * xF;
    const bool xown = L0 && tix < RT_N * xF;
    const bool xok = xown && tile_ok && (tile * WT + xn) < a.B;
    const float *xsrc = nullptr;
    uint8_t *xdst = nullptr;
    float xreg = 0.f;
    const int64_t t_first = dir ? (T - 1) : 0;
    if (L0) {
        const float b0 = a.xin.bias[dir * G3 + j], b1 = a.xin.bias[dir * G3 + H + j], b2 = a.xin.bias[dir * G3 + 2 * H + j];
        br2 = f2_make(b0, b0);
        bz2 = f2_make(b1, b1);
        bn2 = f2_make(b2, b2);
        if (xown) {
            xsrc = a.xin.feats + ((tile * WT + xn) * T) * a.xin.F + xf;
            xdst = smem + L::x_off + X * 2 * RT_XBUF + (xf >> 3) * RT_KG + xn * 16 + (xf & 7) * 2;
            const float x0 = xok ? xsrc[t_first * (int64_t)a.xin.F] : 0.f;
            __half hi, lo;
            split_f16(x0, hi, lo);
            *reinterpret_cast<__half *>(xdst) = hi;
            *reinterpret_cast<__half *>(xdst + RT_XPLANE) = lo;
            if (T > 1 && xok) xreg = xsrc[(dir ? (T - 2) : 1) * (int64_t)a.xin.F];
        }
    }
    const float *xnext = xown ? xsrc + (dir ? (T - 3) : 2) * (int64_t)a.xin.F : nullptr;
    const int64_t xadv = dir ? -(int64_t)a.xin.F : (int64_t)a.xin.F;
    // layer 1: this thread's pre-activations in a staged block: [gate 3][window quad 4][j 128][4 floats]
    const uint8_t *gi_buf = smem + L::gi_off + X * PP_GI_BUFS * PP_GI_BLOCK + ((ch * 2) * H + j) * 16;
    int gbuf = 0;

    F2 hprev2[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) hprev2[p] = f2_make(0.f, 0.f);

    // H_X: 8 gate warps + 2 issuers; N_X: 8 gate warps + relay; both + the tile's copy warp when tiles go out
    constexpr int NB_H = L::OUT_LOG ? PP_NB2 : PP_NB3;
    constexpr int NB_HN = L::OUT_LOG ? PP_NB : PP_NB2;
    // h_{-1} = 0 (zeroed tile) and x_0 are in shared memory: publish
    fence_proxy_async_smem();
    tc_fence_before_sync();
    named_bar_arrive_id<NB_H>(PP_BAR_H + X);

#pragma unroll 1
    for (int64_t s = 0; s < T; ++s) {
        unsigned long long *tr = pp_trace_row<TRACE>(a, s, w8 == 0 && lane == 0);
        const float4 *gs = reinterpret_cast<const float4 *>(gi_buf + gbuf * PP_GI_BLOCK);
        F2 r2[NP], gp2[NP], nzb2[NP], a2[NP], b2[NP];
        // ---------------- r ----------------
        {
            uint32_t ar[8], ax[8];
            F2 gr[NP], gn[NP];
            named_bar_sync_id<PP_NB>(PP_BAR_R + X);
            tc_fence_after_sync();
            tmem_ld_x8(t_lane + 0, ar);
            if (L0) tmem_ld_x8(t_lane + 48, ax);
            if (!L0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float4 vr = gs[(0 * 4 + c) * H], vn = gs[(2 * 4 + c) * H];
                    gr[2 * c] = f2_make(vr.x, vr.y);
                    gr[2 * c + 1] = f2_make(vr.z, vr.w);
                    gn[2 * c] = f2_make(vn.x, vn.y);
                    gn[2 * c + 1] = f2_make(vn.z, vn.w);
                }
            }
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elected) mbar_arrive(&freed[0]);
            PP_STAMP(X * 20 + 7);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (PP_DBG(4)) { r2[p] = one2; gp2[p] = one2; continue; }
                // (weights and biases carry the -log2 e / 2 log2 e factors, common.cuh gate_scale: the accumulator is
                // the exponent.)  r = 1 / (1 + 2^acc), reciprocal on the FMA pipe (rec_common.cuh)
                const F2 accr = f2_make(__uint_as_float(ar[2 * p]), __uint_as_float(ar[2 * p + 1]));
                float e0, e1;
                f2_get(f2_add(L0 ? br2 : gr[p], accr), e0, e1);
                const F2 ea = f2_make(ex2_approx(fminf(e0, EXP_CLAMP)), ex2_approx(fminf(e1, EXP_CLAMP)));
                r2[p] = rcp_neg_fma2(f2_fma(ea, negone2, negone2), one2);
                // the part of the n pre-activation that does not need the n accumulator: gi_n + r * b_hn
                const F2 gin = L0 ? f2_add(bn2, f2_make(__uint_as_float(ax[2 * p]), __uint_as_float(ax[2 * p + 1]))) : gn[p];
                gp2[p] = f2_fma(r2[p], bhn2, gin);
            }
            PP_STAMP(X * 20 + 8);
            phase_fence(gp2[0], gp2[NP - 1]);
        }
        // ---------------- z ----------------
        {
            uint32_t az[8];
            F2 gz[NP];
            named_bar_sync_id<PP_NB>(PP_BAR_Z + X);
            tc_fence_after_sync();
            tmem_ld_x8(t_lane + 16, az);
            if (!L0) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const float4 vz = gs[(1 * 4 + c) * H];
                    gz[2 * c] = f2_make(vz.x, vz.y);
                    gz[2 * c + 1] = f2_make(vz.z, vz.w);
                }
            }
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elected) mbar_arrive(&freed[1]);
            PP_STAMP(X * 20 + 9);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (PP_DBG(4)) { nzb2[p] = negone2; a2[p] = one2; b2[p] = one2; continue; }
                // z = 1 / (1 + eb) is never formed.  With en = e^{2x} (the tanh argument, next phase):
                //   h = (1 - z) n + z h_prev = ((eb + h_prev) en + (h_prev - eb)) / ((1 + eb) en + (1 + eb))
                // so everything but en is prepared here, off the n -> h critical path
                const F2 accz = f2_make(__uint_as_float(az[2 * p]), __uint_as_float(az[2 * p + 1]));
                float e0, e1;
                f2_get(f2_add(L0 ? bz2 : gz[p], accz), e0, e1);
                const F2 eb = f2_make(ex2_approx(fminf(e0, EXP_CLAMP)), ex2_approx(fminf(e1, EXP_CLAMP)));
                nzb2[p] = f2_fma(eb, negone2, negone2);              // -(1 + eb)
                a2[p] = f2_add(eb, hprev2[p]);
                b2[p] = f2_fma(eb, negone2, hprev2[p]);
            }
            PP_STAMP(X * 20 + 10);
            phase_fence(a2[0], b2[NP - 1]);
        }
        // ---------------- n, h ----------------
        {
            uint32_t an[8];
            named_bar_sync_id<NB_HN>(PP_BAR_N + X);
            tc_fence_after_sync();
            tmem_ld_x8(t_lane + 32, an);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (elected) mbar_arrive(&freed[2]);
            PP_STAMP(X * 20 + 11);
            uint8_t *hw = hrow + (int)((s + 1) & 1) * (2 * RT_HPLANE);      // the buffer the MMAs of this step do not read
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (PP_DBG(4)) continue;
                const F2 accn = f2_make(__uint_as_float(an[2 * p]), __uint_as_float(an[2 * p + 1]));
                float t0, t1;
                f2_get(f2_fma(r2[p], accn, gp2[p]), t0, t1);
                const F2 en = f2_make(ex2_approx(fminf(t0, EXP_CLAMP)), ex2_approx(fminf(t1, EXP_CLAMP)));
                // exponentials are clamped at 2^60, so numerator and denominator stay below 2^122
                const F2 inv = rcp_neg_fma2(f2_fma(nzb2[p], en, nzb2[p]), one2);
                const F2 h2 = f2_mul(f2_fma(a2[p], en, b2[p]), inv);
                hprev2[p] = h2;
                // fp16 hi/lo split by truncation: hi = h with the low 13 mantissa bits cleared (exactly an fp16 value for
                // |h| >= 2^-14), lo = fp16(h - hi): 22+ significant bits (rec_tc_kernel has the details)
                float h0v, h1v;
                f2_get(h2, h0v, h1v);
                const float u0 = __uint_as_float(__float_as_uint(h0v) & 0xFFFFE000u);
                const float u1 = __uint_as_float(__float_as_uint(h1v) & 0xFFFFE000u);
                float l0v, l1v;
                f2_get(f2_fma(f2_make(u0, u1), negone2, h2), l0v, l1v);
                const __half2 hi2 = __floats2half2_rn(u0, u1), lo2 = __floats2half2_rn(l0v, l1v);
                if (PP_DBG(1)) continue;
                *reinterpret_cast<__half *>(hw + (2 * p) * 16) = __low2half(hi2);
                *reinterpret_cast<__half *>(hw + RT_HPLANE + (2 * p) * 16) = __low2half(lo2);
                *reinterpret_cast<__half *>(hw + (2 * p + 1) * 16) = __high2half(hi2);
                *reinterpret_cast<__half *>(hw + RT_HPLANE + (2 * p + 1) * 16) = __high2half(lo2);
            }
            if (L0 && xown && s + 1 < T && !PP_DBG(8)) {
                // stage x_{s+1} (loaded a step ago) into the other x buffer
                __half hi, lo;
                split_f16(xreg, hi, lo);
                uint8_t *xd = xdst + (((s + 1) & 1) ? RT_XBUF : 0);
                *reinterpret_cast<__half *>(xd) = hi;
                *reinterpret_cast<__half *>(xd + RT_XPLANE) = lo;
            }
            PP_STAMP(X * 20 + 12);
            if (!PP_DBG(2))
            fence_proxy_async_smem();     // h / x tile writes -> visible to the MMAs' async-proxy reads
            tc_fence_before_sync();
            if (s + 1 < T) named_bar_arrive_id<NB_H>(PP_BAR_H + X);
            else named_bar_arrive_id<PP_NB>(PP_BAR_FIN + X);
            PP_STAMP(X * 20 + 13);
        }
        if (L0 && xok && s + 2 < T && !PP_DBG(8)) {
            xreg = *xnext;                // the feature value staged during the NEXT step
            xnext += xadv;
        }
        gbuf = (gbuf == PP_GI_BUFS - 1) ? 0 : gbuf + 1;
    }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int LAYER, bool ALLP, bool TRACE>
__global__ void __launch_bounds__(PP_THREADS, 1) rec_pp_kernel(const __grid_constant__ PPArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    using L = PPCfg<LAYER>;
    constexpr bool IN_X = L::IN_X, OUT_LOG = L::OUT_LOG;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::bar_off);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::tmem_off);
    const int tid = threadIdx.x, lane = tid & 31;
    // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so that everything derived from it -
    // the tile index, shared-memory descriptors - can live in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int dir = blockIdx.y;

    // ---- prologue: zero the tiles (h_{-1} = 0; empty tiles compute on zeros), barriers, TMEM, weights ----
    {
        int4 *z = reinterpret_cast<int4 *>(smem);
        for (int i = tid; i < L::bar_off / 16; i += PP_THREADS) z[i] = make_int4(0, 0, 0, 0);
    }
    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);              // acc[X][g]: one tcgen05.commit each
        for (int i = 6; i < 12; ++i) mbar_init(&bars[i], 1);             // gi_full[X][b]: expect_tx + bytes
        for (int i = 12; i < 18; ++i) mbar_init(&bars[i], PP_TILE_WARPS);   // free_acc[X][g]: one lane per gate warp of a tile
        mbar_init(&bars[18], 1);                                         // log_full: commit
        mbar_init(&bars[19], 1);                                         // log_free[X]: the logits warp
        mbar_init(&bars[20], 1);
        fence_mbar_init();
    }
    if (warp == PP_W_AUX0) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    __syncthreads();     // zeroing done before anything else is written to shared memory
    if (OUT_LOG) {       // this direction's half of W_lin, lo plane, already in operand layout (pack_linear_kernel)
        const int4 *src = reinterpret_cast<const int4 *>(a.lin_w_tc + ((size_t)dir * 2 + 1) * (PP_WL_PLANE / 2));
        int4 *dst = reinterpret_cast<int4 *>(smem + L::wl_off);
        for (int i = tid; i < PP_WL_PLANE / 16; i += PP_THREADS) dst[i] = src[i];
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0u) {   // a 512-column allocation starts at column 0; the issuers rely on literal addresses
        if (tid == 0) printf("mdk: unexpected TMEM base %u for a 512-column allocation\n", tmem_base);
        __trap();
    }
    // weights (row-major fp16 hi/lo) -> TMEM: lane j, 8 columns per K = 16 chunk, cell = (k even | k odd << 16)
    if (warp < 4) {
        const int jrow = warp * 32 + lane;
        const uint32_t t_w = (uint32_t)(warp * 32) << 16;
        for (int pg = 0; pg < 6; ++pg) {   // pg = part*3 + gate
            const uint4 *src = reinterpret_cast<const uint4 *>(a.w_hh + (((size_t)dir * 6 + pg) * H + jrow) * H);
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                const uint4 lo4 = src[2 * ks], hi4 = src[2 * ks + 1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + (uint32_t)((pg * 8 + ks) * 8), v);
            }
            if (IN_X) {
                const uint4 *sx = reinterpret_cast<const uint4 *>(a.xin.w_x + (((size_t)dir * 6 + pg) * H + jrow) * 16);
                const uint4 lo4 = sx[0], hi4 = sx[1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + L::wx_col + (uint32_t)(pg * 8), v);
            }
        }
        if (OUT_LOG) {   // W_lin hi plane (rows >= 5 zero), the row-major copy behind the two shared-memory images
            const uint4 *sl = reinterpret_cast<const uint4 *>(a.lin_w_tc + (size_t)NDIR * 2 * (PP_WL_PLANE / 2) +
                                                              ((size_t)dir * H + jrow) * H);
#pragma unroll
            for (int ks = 0; ks < H / 16; ++ks) {
                const uint4 lo4 = sl[2 * ks], hi4 = sl[2 * ks + 1];
                const uint32_t v[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                tmem_st_x8(t_w + L::wl_col + (uint32_t)(ks * 8), v);
            }
        }
        tmem_st_wait();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();

    // the two tiles run the same code with the tile index in a register (one copy in the instruction cache)
    if (warp < PP_GATE_WARPS) pp_gate<LAYER, TRACE>(smem, a, warp / PP_TILE_WARPS, warp % PP_TILE_WARPS, lane);
    else if (warp >= PP_W_ISS && warp < PP_W_ISS + 4) pp_issuer<LAYER, ALLP, TRACE>(smem, a, (warp - PP_W_ISS) & 1, (warp - PP_W_ISS) >> 1);
    else if (warp == PP_W_REL || warp == PP_W_REL + 1) pp_relay<LAYER, TRACE>(smem, a, warp - PP_W_REL, lane);
    else if (OUT_LOG) {
        if (warp == PP_W_AUX0) pp_logits<TRACE>(smem, a, lane);
    } else {
        pp_copy<LAYER, TRACE>(smem, a, warp == PP_W_AUX0 ? 0 : 1, lane);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == PP_W_AUX0) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 512);
    }
}

unsigned long long *rec_trace_buffer();   // gru_tc.cu: [2 layers][steps][slots], null = tracing off
static uint32_t g_pp_debug = 0;
void pp_set_debug(uint32_t flags) { g_pp_debug = flags; }

cudaError_t launch_rec_pp(int layer, const float *gi, const RecXArgs *fuse, const __half *w_hh_tm, const float *b_hn,
                          void *h_out, int64_t B, int64_t T, cudaStream_t s, const __half *lin_w_tc, float *plog,
                          uint32_t prod_mask) {
    if (B == 0 || T == 0) return cudaSuccess;
    PPArgs a{};
    a.gi = gi;
    if (fuse) a.xin = RecX{fuse->feats, fuse->w_x, fuse->bias, fuse->F};
    a.w_hh = w_hh_tm;
    a.b_hn = b_hn;
    a.h_out = h_out;
    a.B = B;
    a.T = T;
    a.ntiles = (B + RT_N - 1) / RT_N;
    a.lin_w_tc = lin_w_tc;
    a.plog = plog;
    a.prod_mask = prod_mask & 7u;
    a.debug = g_pp_debug;
    // layer: 0 = layer 0 (fused projection when `fuse` is given, else gi in), 1 = layer 1
    const int variant = layer == 1 ? 1 : (fuse ? 0 : 2);
    if (variant == 0 && (fuse->F > 16 || !h_out)) return cudaErrorInvalidValue;
    if (variant == 1 && (!gi || !lin_w_tc || !plog)) return cudaErrorInvalidValue;
    if (variant == 2 && (!gi || !h_out)) return cudaErrorInvalidValue;
    if ((a.prod_mask & 1u) == 0) return cudaErrorInvalidValue;   // the hi x hi product is not optional
    unsigned long long *tb = rec_trace_buffer();
    const bool trace = tb && T >= PP_TRACE_STEP0 + PP_TRACE_STEPS;
    if (trace) a.trace = tb + layer * PP_TRACE_STEPS * PP_TRACE_SLOTS;
    const dim3 grid((unsigned)((a.ntiles + 1) / 2), NDIR);
    cudaError_t e;
#define MDK_LAUNCH_PP(LY, AP, TR)                                                                          \
    do {                                                                                                   \
        auto kern = rec_pp_kernel<LY, AP, TR>;                                                             \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PPCfg<LY>::total);     \
        if (e != cudaSuccess) return e;                                                                    \
        kern<<<grid, PP_THREADS, PPCfg<LY>::total, s>>>(a);                                                \
    } while (0)
    const bool allp = a.prod_mask == 7u;
    if (variant == 0) {
        if (trace) MDK_LAUNCH_PP(0, true, true);
        else if (allp) MDK_LAUNCH_PP(0, true, false);
        else MDK_LAUNCH_PP(0, false, false);
    } else if (variant == 1) {
        if (trace) MDK_LAUNCH_PP(1, true, true);
        else if (allp) MDK_LAUNCH_PP(1, true, false);
        else MDK_LAUNCH_PP(1, false, false);
    } else {
        if (allp) MDK_LAUNCH_PP(2, true, false);
        else MDK_LAUNCH_PP(2, false, false);
    }
#undef MDK_LAUNCH_PP
    return cudaGetLastError();
}

}  // namespace mdk
