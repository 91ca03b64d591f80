// Shared declarations of libmedaka_b200: engine state, error plumbing, kernel launchers.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/medaka_b200.h"

namespace mdk {

constexpr int H = 128;        // gru_size of every shipped counts-matrix model (gru.py:18)
constexpr int G3 = 3 * H;     // gate rows per direction (r,z,n)
constexpr int NDIR = 2;
constexpr int GI_COLS = NDIR * G3;  // 768: input-projection row [fwd r z n | rev r z n]
constexpr int H2 = NDIR * H;        // 256: layer output width
constexpr int NCLS = 5;             // gru.py:53-55

void set_error(const std::string &msg);
int cuda_fail(cudaError_t err, const char *what, const char *file, int line);

#define MDK_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return ::mdk::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define MDK_REQUIRE(cond, code, msg)  \
    do {                              \
        if (!(cond)) {                \
            ::mdk::set_error(msg);    \
            return (code);            \
        }                             \
    } while (0)

// Geometry of the fp16 hi/lo operand tiles the tcgen05 kernels consume (K-major, no swizzle:
// [k-group][row][8 halfs]; see ptx.cuh make_smem_desc).
constexpr int XT_ROWS = 128;                       // positions per activation tile
constexpr int XT_K = H2;                           // 256
constexpr int XT_PLANE_BYTES = XT_ROWS * XT_K * 2; // one plane (hi or lo) of a tile: 64 KiB
constexpr int XT_TILE_BYTES = 2 * XT_PLANE_BYTES;  // hi plane then lo plane

// Tile-interleaved row order of the tensor-core path's intermediates (gi, h0 tiles, h1): windows are grouped in
// tiles of WT = 16 (the recurrent kernel's N) and rows run  p' = ((w / 16) * T + t) * 16 + w % 16 , i.e. the 16
// windows of a tile are adjacent for a given time step.  A thread's windows are then `base + c * const` apart
// (immediate offsets, no per-element address arithmetic, no ragged-edge predicates: padding windows are real rows
// that are simply never copied out), and one tile-step of gi is a single contiguous 48 KiB block.
constexpr int WT = 16;
__host__ __device__ inline int64_t tiled_rows(int64_t B, int64_t T) { return ((B + WT - 1) / WT) * WT * T; }
__host__ __device__ inline int64_t tiled_row(int64_t w, int64_t t, int64_t T) { return ((w / WT) * T + t) * WT + (w % WT); }

// gi (gate pre-activations) on the tensor-core path is stored in QUAD layout: per tile-step ts = tiled row / 16 (the 16
// windows of a tile at one time step) a 48 KiB block  [blk = dir*3 + gate (6)][cg = window quad (4)][j (128)][w4 (4)]
// of floats.  A recurrent-kernel thread (hidden unit j, window quad cg) then reads its 3 gates x 4 windows as three
// 16-byte loads that are contiguous across the warp (512 B per load), the projection GEMM's epilogue thread (row j of
// block blk) writes 16-byte vectors that are contiguous across its warp, and the (ts, dir) block a CTA needs next is
// one contiguous 24 KiB range for the L2 prefetch.
constexpr int GI_TS_FLOATS = 6 * 4 * H * 4;   // 12288 floats per tile-step
__host__ __device__ inline int64_t gi_quad_index(int64_t tiled_row, int col) {
    const int64_t ts = tiled_row >> 4;
    const int w = (int)(tiled_row & 15), blk = col / H, j = col % H;
    return ((((ts * 6 + blk) * 4 + (w >> 2)) * H + j) << 2) + (w & 3);
}

// sigmoid(a) = 1 / (1 + 2^(-log2e * a)),  tanh(x) = (2^(2 log2e * x) - 1) / (2^(2 log2e * x) + 1)
constexpr float GATE_SCALE_RZ = -1.4426950408889634f;
constexpr float GATE_SCALE_N = 2.8853900817779268f;
__host__ __device__ inline float gate_scale(int gate) { return gate < 2 ? GATE_SCALE_RZ : GATE_SCALE_N; }

struct LayerWeights {
    // fp32 originals (device), torch layout
    float *w_ih[NDIR] = {nullptr, nullptr};  // [3H][in]
    float *w_hh[NDIR] = {nullptr, nullptr};  // [3H][H]
    float *b_ih[NDIR] = {nullptr, nullptr};
    float *b_hh[NDIR] = {nullptr, nullptr};
    bool loaded[NDIR] = {false, false};
    // derived (built by prepare_weights)
    float *w_in_packed = nullptr;   // [768][in] fp32: rows = dir*384 + gate*128 + j
    float *bias_gi = nullptr;       // [768]: r,z: b_ih+b_hh ; n: b_ih
    float *b_hn = nullptr;          // [2][128]
    // tensor-core path: the activations' exp2 scale factors are folded into everything that feeds a gate pre-activation
    // (rows of W_hh / W_ih and the biases of the r and z gates times -log2 e, of the n gate times 2 log2 e), so the gate
    // math goes straight from the accumulators into ex2 - see gate_scale() and rec_tc_kernel
    float *bias_gi_tc = nullptr;    // [768] = bias_gi * gate_scale
    float *b_hn_tc = nullptr;       // [2][128] = b_hn * gate_scale(n)
    float *w_hh_t = nullptr;        // [2][128(k)][384] fp32 (transposed) for the FFMA path
    __half *w_hh_tm = nullptr;      // [2][hi/lo][gate][row128][k128] fp16 row-major: source of the TMEM-resident A operand
    __half *w_x_tm = nullptr;       // layer 0, F <= 16: [2][hi/lo][gate][row128][16] fp16 (K zero-padded): fused input projection
    __half *w_in_tc = nullptr;      // layer 1 only: [6 blocks][hi/lo][row128][k256] fp16 row-major: source of the
                                    // gemm_tc TMEM-resident A operand
};

}  // namespace mdk

// The engine runs GROUPS of windows.  A workspace (mdk_ws) is a compute stream plus the intermediates of one forward; a
// lane (mdk_lane) is the device-side staging of one group of submitted batches (features in, probabilities / labels
// out) and is bound to one workspace.  There are more big lanes than big workspaces: while two groups compute, a third
// is receiving its features and a fourth is draining its results, so the copies never hold a workspace (48 GB for a
// 1184 x 10 000 group) idle.  Groups on the two big workspaces run concurrently: the ping-pong recurrent kernels
// (gru_pp.cu) need half of the SMs for a 1184-window group.  Small forwards (the B = 1 remainder regions of
// medaka/prediction.py:196-209) spread over the small lanes, each with a workspace of its own.
struct mdk_ws {
    cudaStream_t stream = nullptr;
    int64_t cap_pos = 0;       // capacity in positions (rounded up to XT_ROWS)
    float *gi = nullptr;       // [cap_pos][768]
    void *h0 = nullptr;        // fp32 [cap_pos][256]  or  fp16 hi/lo tiles (same byte size)
    float *h1 = nullptr;       // [cap_pos][256], allocated on first use (unfused head only)
    int64_t cap_h1 = 0;
    float *plog = nullptr;     // fused head: per-direction partial logits [dir][tile-step][class 5][16 windows]
    int *gemm_ctr = nullptr;   // the GEMM's six tile counters (zeroed in front of each launch)
    // geometry of the last forward run here (mdk_engine_read_activation)
    int64_t last_B = 0, last_T = 0;
    int last_precision = -1;
    bool last_fused_head = false;
};

struct mdk_lane {
    mdk_ws *ws = nullptr;      // where the lane's groups compute (fixed at engine creation)
    // device staging of the group's host buffers
    int64_t cap_io = 0;        // positions
    int64_t cap_feats = 0;     // floats
    float *d_feats = nullptr, *d_probs = nullptr, *d_logits = nullptr;
    uint8_t *d_labels = nullptr;
    cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
    // the group being collected (open) or in flight (busy)
    struct Item {
        const float *feats;
        float *probs, *logits;
        uint8_t *labels;
        int64_t B;
    };
    std::vector<Item> items;
    int64_t gB = 0, gT = 0;
    bool want_logits = false, want_labels = false;
    bool open = false, busy = false;
    int64_t group = -1;        // serial number of the lane's current (open / in-flight / last finished) group
};

struct mdk_engine {
    int device = 0;
    mdk_model_desc desc{};
    int precision = MDK_PREC_TC;
    int sm_count = 148;
    bool fuse_x = true;           // layer-0 input projection fused into the recurrence (F <= 16); MDK_NO_FUSE_X=1 disables
    int rec_mode = MDK_REC_AUTO;  // which recurrent kernel the tensor-core path runs (MDK_REC_*)
    uint32_t prod_mask = 7u;      // fp16 products per contraction (mdk_engine_set_products)
    static constexpr int BIG_WS = 2, BIG_LANES = 4, SMALL_LANES = 14;
    static constexpr int N_LANES = BIG_LANES + SMALL_LANES, N_WS = BIG_WS + SMALL_LANES;
    static constexpr int64_t SMALL_POS = 1 << 18;   // forwards up to this many positions run on the small lanes
    mdk_ws ws[N_WS];              // [0, BIG_WS): big groups; then one per small lane
    mdk_lane lane[N_LANES];       // big lane j computes on ws[j % BIG_WS], small lane i on ws[BIG_WS + i]
    int next_big = 0, next_small = 0, next_big_ws = 0;
    int open_lane = -1;           // lane whose group is still collecting batches (at most one), -1 = none
    int last_ws = 0;              // workspace of the most recent forward
    int64_t group_windows = 0;    // most windows coalesced into one group (0 = one wave, mdk_engine_preferred_windows)
    cudaStream_t stream = nullptr;       // == ws[0].stream: weight preparation, timers
    static constexpr int EV_RING = 32;   // per-forward event sets kept for mdk_engine_mean_timings
    cudaEvent_t evr[EV_RING][8] = {};
    cudaEvent_t *ev = evr[0];            // event set of the forward being queued
    int64_t fwd_count = 0;
    cudaEvent_t ev_timer[2] = {};
    cudaEvent_t ev_join = nullptr;
    mdk::LayerWeights layer[2];
    float *lin_w = nullptr, *lin_b = nullptr;
    bool lin_loaded = false;
    __half *lin_w_tc = nullptr;   // [dir][hi/lo][k-group 16][row 64][8 halfs]: W_lin half of one direction as an M=64 smem A
                                  // operand (rows >= 5 zero) for the logits MMAs fused into the layer-1 recurrence
    bool keep_act = false;        // debugging: keep h1 (layer-1 output) in HBM, i.e. run the unfused head
    bool prepared = false;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // tickets: ticket -> (lane, group) for the last TICKET_RING submits; older ones have completed (their lane was reused)
    static constexpr int TICKET_RING = 4096;
    int16_t ticket_lane[TICKET_RING] = {};
    int64_t ticket_group[TICKET_RING] = {};
    int64_t submit_count = 0;
    mdk_timings last{};
    int64_t launches = 0;
};

namespace mdk {

// ---- launchers (each returns cudaGetLastError() of its launch) -------------------------------
// misc.cu
// tiled != 0: gi rows are written / h1 rows are read in tile-interleaved order (T = window length)
cudaError_t launch_inproj0(const float *feats, const float *w_packed, const float *bias, float *gi,
                           int64_t P, int F, int64_t T, int tiled, cudaStream_t s);
cudaError_t launch_head(const float *h1, const float *lin_w, const float *lin_b, int64_t B, int64_t T, int tiled,
                        float *probs, float *logits, uint8_t *labels, cudaStream_t s);
cudaError_t launch_untile_rows(const float *src_tiled, float *dst, int64_t B, int64_t T, cudaStream_t s);
cudaError_t launch_normalise(const uint64_t *counts, const int64_t *major, const int64_t *minor, int64_t n,
                             int num_dtypes, int mode, int sym_indels, float *feats, int64_t *depth,
                             cudaStream_t s);
cudaError_t launch_decode(const float *probs, int64_t n, uint8_t *labels, uint8_t *quals, cudaStream_t s);
cudaError_t launch_decode_f64(const double *probs, int64_t n, uint8_t *labels, uint8_t *quals, cudaStream_t s);
cudaError_t launch_variant_columns(const int64_t *minor, const uint8_t *ref, const uint8_t *pred, int64_t n,
                                   uint8_t *out, cudaStream_t s);
cudaError_t launch_prepare_layer(const LayerWeights &lw, int in_features, bool build_in_tc, cudaStream_t s);
cudaError_t launch_unpack_h0(const void *h0_tiles, float *out, int64_t B, int64_t T, cudaStream_t s);
// gru_fp32.cu
cudaError_t launch_rec_fp32(const float *gi, const float *w_hh_t, const float *b_hn, float *h_out, int64_t B,
                            int64_t T, cudaStream_t s);
cudaError_t launch_gemm_fp32(const float *A, const float *W, const float *bias, float *C, int64_t P,
                             cudaStream_t s);
// gru_tc.cu
struct RecXArgs {            // fused layer-0 input projection (rec_tc FUSE_X)
    const float *feats;      // [B][T][F]
    const __half *w_x;       // LayerWeights::w_x_tm
    const float *bias;       // LayerWeights::bias_gi
    int F;
};
cudaError_t rec_trace_control(int enable, unsigned long long *host_out);
// lin_w_tc != nullptr (layer 1, one tile per CTA): the 5-class linear head runs inside the recurrence as extra MMAs and
// the kernel writes partial logits to plog instead of h_out; *fused_logits tells the caller whether it did
bool rec_tc_can_fuse_logits(int64_t B, int sm_count);
cudaError_t launch_rec_tc(const float *gi, const RecXArgs *fuse, const __half *w_hh_tm, const float *b_hn,
                          void *h_out, int out_tiles, int64_t B, int64_t T, int sm_count, cudaStream_t s,
                          const __half *lin_w_tc = nullptr, float *plog = nullptr, uint32_t prod_mask = 7u);
// gru_pp.cu: two tiles per CTA (ping-pong).  layer 0: fused projection when `fuse` is given (F <= 16), else gi in; operand
// tiles out.  layer 1: gi in, partial logits out (lin_w_tc, plog required).  prod_mask: fp16 products per contraction
// (bit 0 W_hi.h_hi, bit 1 W_hi.h_lo, bit 2 W_lo.h_hi; 7 = fp32-faithful)
cudaError_t launch_rec_pp(int layer, const float *gi, const RecXArgs *fuse, const __half *w_hh_tm, const float *b_hn,
                          void *h_out, int64_t B, int64_t T, cudaStream_t s, const __half *lin_w_tc, float *plog,
                          uint32_t prod_mask);
void pp_set_debug(uint32_t flags);   // diagnostics of the traced ping-pong kernels (gru_pp.cu PPArgs::debug)
// head on the partial logits of the fused path: sum of the two directions + bias -> softmax / argmax
cudaError_t launch_head_plog(const float *plog, const float *lin_b, int64_t B, int64_t T, float *probs, float *logits,
                             uint8_t *labels, cudaStream_t s);
cudaError_t launch_pack_linear(const float *lin_w, __half *lin_w_tc, cudaStream_t s);
constexpr int PLOG_TS_FLOATS = NCLS * WT;     // 80 floats per (tile-step, direction)
cudaError_t launch_gemm_tc(const void *x_tiles, const __half *w_in_tm, const float *bias, float *gi, int64_t P,
                           int sm_count, cudaStream_t s, uint32_t prod_mask, int *tile_ctr);
int selftest_umma(int device, const float *A, const float *B, float *D, int N, int K, int variant);
// pileup.cu
cudaError_t plp_scratch(size_t bytes, uint8_t **out, int slot);   // per-host-thread cached device buffers (slot 0 / 1)
int pileup_counts_dev(int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                      const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, int64_t n_ops,
                      const uint8_t *seq, const int64_t *seq_off, int32_t start, int32_t end, int num_dtypes,
                      int min_mapq, int64_t max_cols, uint64_t *counts, int64_t *major, int64_t *minor,
                      int64_t *n_cols_host, cudaStream_t s);

}  // namespace mdk
