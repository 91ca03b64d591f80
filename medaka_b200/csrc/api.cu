// C ABI of libmedaka_b200 (include/medaka_b200.h): engine life-cycle, weight loading, the forward
// pipeline, and the featuriser / decode entry points.  Host orchestration only - kernels live in
// misc.cu, gru_fp32.cu and gru_tc.cu.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"

namespace mdk {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t err, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)err, cudaGetErrorString(err), file, line, what);
    g_last_error = buf;
    if (err == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        return MDK_ERR_NOMEM;
    }
    return MDK_ERR_CUDA;
}

template <typename T>
static int dev_alloc(T **p, size_t n_elems) {
    MDK_CUDA(cudaMalloc(reinterpret_cast<void **>(p), n_elems * sizeof(T)));
    return MDK_OK;
}
template <typename T>
static void dev_free(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// Host -> device upload ordered on the ENGINE stream.  A plain cudaMemcpy from pageable memory returns once the
// data is staged, possibly before the DMA lands, and the engine stream is non-blocking (it does not synchronise
// with the legacy default stream): the packing kernels could then read stale weights (seen as run-to-run
// 3e-6 differences in layer-0 outputs).  cudaMemcpyAsync + stream sync closes that window.
static int upload(cudaStream_t stream, float **dst, const float *src, size_t n) {
    if (!*dst) {
        int rc = dev_alloc(dst, n);
        if (rc) return rc;
    }
    MDK_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    MDK_CUDA(cudaStreamSynchronize(stream));
    return MDK_OK;
}

static int in_features(const mdk_engine *e, int layer) { return layer == 0 ? e->desc.num_features : H2; }

static int prepare_weights(mdk_engine *e) {
    if (e->prepared) return MDK_OK;
    for (int l = 0; l < 2; ++l) {
        MDK_REQUIRE(e->layer[l].loaded[0] && e->layer[l].loaded[1], MDK_ERR_STATE,
                    "engine: GRU weights not loaded for every (layer, direction)");
    }
    MDK_REQUIRE(e->lin_loaded, MDK_ERR_STATE, "engine: linear head weights not loaded");
    for (int l = 0; l < 2; ++l) {
        LayerWeights &lw = e->layer[l];
        const int in = in_features(e, l);
        int rc;
        if (!lw.w_in_packed && (rc = dev_alloc(&lw.w_in_packed, (size_t)GI_COLS * in))) return rc;
        if (!lw.bias_gi && (rc = dev_alloc(&lw.bias_gi, (size_t)GI_COLS))) return rc;
        if (!lw.b_hn && (rc = dev_alloc(&lw.b_hn, (size_t)NDIR * H))) return rc;
        if (!lw.bias_gi_tc && (rc = dev_alloc(&lw.bias_gi_tc, (size_t)GI_COLS))) return rc;
        if (!lw.b_hn_tc && (rc = dev_alloc(&lw.b_hn_tc, (size_t)NDIR * H))) return rc;
        if (!lw.w_hh_t && (rc = dev_alloc(&lw.w_hh_t, (size_t)NDIR * H * G3))) return rc;
        if (!lw.w_hh_tm && (rc = dev_alloc(&lw.w_hh_tm, (size_t)NDIR * 2 * G3 * H))) return rc;
        if (l == 0 && in <= 16 && !lw.w_x_tm && (rc = dev_alloc(&lw.w_x_tm, (size_t)NDIR * 2 * G3 * 16))) return rc;
        if (l == 1 && !lw.w_in_tc && (rc = dev_alloc(&lw.w_in_tc, (size_t)2 * GI_COLS * H2))) return rc;
        MDK_CUDA(launch_prepare_layer(lw, in, l == 1, e->stream));
        e->launches++;
    }
    {
        int rc;
        if (!e->lin_w_tc && (rc = dev_alloc(&e->lin_w_tc, (size_t)NDIR * 2 * 16 * 64 * 8 + (size_t)NDIR * H * H))) return rc;
        MDK_CUDA(launch_pack_linear(e->lin_w, e->lin_w_tc, e->stream));
        e->launches++;
    }
    MDK_CUDA(cudaStreamSynchronize(e->stream));
    e->prepared = true;
    return MDK_OK;
}

static int ensure_workspace(mdk_engine *e, mdk_ws &ws, int64_t B, int64_t T) {
    // rows of the tile-interleaved intermediates (>= B*T: the last window tile is padded to 16 windows)
    const int64_t rows = tiled_rows(B, T);
    const int64_t need = ((rows + XT_ROWS - 1) / XT_ROWS) * XT_ROWS;
    if (need <= ws.cap_pos) return MDK_OK;
    MDK_CUDA(cudaStreamSynchronize(ws.stream));
    dev_free(ws.gi);
    if (ws.h0) { cudaFree(ws.h0); ws.h0 = nullptr; }
    dev_free(ws.h1);
    ws.cap_h1 = 0;
    dev_free(ws.plog);
    ws.cap_pos = 0;
    int rc;
    if ((rc = dev_alloc(&ws.gi, (size_t)need * GI_COLS))) return rc;
    MDK_CUDA(cudaMalloc(&ws.h0, (size_t)need * H2 * sizeof(float)));
    if ((rc = dev_alloc(&ws.plog, (size_t)NDIR * (need / WT) * PLOG_TS_FLOATS))) return rc;
    if (!ws.gemm_ctr) MDK_CUDA(cudaMalloc(&ws.gemm_ctr, 8 * sizeof(int)));
    ws.cap_pos = need;
    return MDK_OK;
}

// h1 (the layer-1 output, 1 KiB / position) only exists on the unfused-head paths: allocated on first use
static int ensure_h1(mdk_ws &ws) {
    if (ws.cap_h1 >= ws.cap_pos && ws.h1) return MDK_OK;
    MDK_CUDA(cudaStreamSynchronize(ws.stream));
    dev_free(ws.h1);
    ws.cap_h1 = 0;
    int rc;
    if ((rc = dev_alloc(&ws.h1, (size_t)ws.cap_pos * H2))) return rc;
    ws.cap_h1 = ws.cap_pos;
    return MDK_OK;
}

static int ensure_io(mdk_engine *e, mdk_lane &ln, int64_t B, int64_t T) {
    const int64_t P = B * T;
    const int64_t feats = P * e->desc.num_features;
    if (P <= ln.cap_io && feats <= ln.cap_feats) return MDK_OK;
    MDK_CUDA(cudaStreamSynchronize(ln.ws->stream));
    MDK_CUDA(cudaStreamSynchronize(e->copy_in));
    MDK_CUDA(cudaStreamSynchronize(e->copy_out));
    ln.cap_io = 0; ln.cap_feats = 0;
    dev_free(ln.d_feats); dev_free(ln.d_probs); dev_free(ln.d_logits); dev_free(ln.d_labels);
    int rc;
    if ((rc = dev_alloc(&ln.d_feats, (size_t)feats))) return rc;
    if ((rc = dev_alloc(&ln.d_probs, (size_t)P * NCLS))) return rc;
    if ((rc = dev_alloc(&ln.d_logits, (size_t)P * NCLS))) return rc;
    if ((rc = dev_alloc(&ln.d_labels, (size_t)P))) return rc;
    ln.cap_io = P; ln.cap_feats = feats;
    return MDK_OK;
}

// Which recurrent kernel runs a batch of B windows (tensor-core path).  One tile per CTA (rec_tc_kernel, NT = 1) is a
// single dependent chain per SM; two tiles per CTA (rec_pp_kernel) interleave two chains on one SM and need half as many
// CTAs, so two groups on two lanes share the GPU.  AUTO: ping-pong from half a wave of tiles up.
static bool use_pingpong(const mdk_engine *e, int64_t B) {
    const int64_t tiles = (B + WT - 1) / WT;
    if (e->rec_mode == MDK_REC_PINGPONG) return true;
    if (e->rec_mode == MDK_REC_ONE_TILE) return false;
    return tiles * NDIR > (int64_t)e->sm_count / 2;
}

// The forward pipeline on the workspace's stream.  ev[1..6] bracket the stages for mdk_timings.
static int run_forward(mdk_engine *e, mdk_ws &ws, const float *feats_dev, int64_t B, int64_t T, float *probs_dev,
                       float *logits_dev, uint8_t *labels_dev) {
    int rc;
    if ((rc = prepare_weights(e))) return rc;
    if ((rc = ensure_workspace(e, ws, B, T))) return rc;
    const int64_t P = B * T;
    cudaStream_t s = ws.stream;
    const bool tc = e->precision == MDK_PREC_TC;
    int launches = 0;
    const bool fuse_x = tc && e->fuse_x && e->layer[0].w_x_tm != nullptr;
    const bool pp = tc && use_pingpong(e, B);
    // the linear head rides inside the layer-1 recurrence as extra MMAs (40 B/position of partial logits reach HBM
    // instead of the 1 KiB/position h1 round trip): always on the ping-pong path, on the one-tile path when the batch
    // is one tile per CTA
    const bool fuse_head = tc && !e->keep_act && (pp || rec_tc_can_fuse_logits(B, e->sm_count));
    if (!fuse_head && (rc = ensure_h1(ws))) return rc;
    MDK_CUDA(cudaEventRecord(e->ev[1], s));
    if (!fuse_x) {
        MDK_CUDA(launch_inproj0(feats_dev, e->layer[0].w_in_packed, e->layer[0].bias_gi, ws.gi, P,
                                e->desc.num_features, T, tc ? 1 : 0, s));
        launches++;
    }
    MDK_CUDA(cudaEventRecord(e->ev[2], s));
    if (tc) {
        const RecXArgs fx{feats_dev, e->layer[0].w_x_tm, e->layer[0].bias_gi_tc, e->desc.num_features};
        if (pp) MDK_CUDA(launch_rec_pp(0, ws.gi, fuse_x ? &fx : nullptr, e->layer[0].w_hh_tm, e->layer[0].b_hn_tc, ws.h0, B, T, s,
                                       nullptr, nullptr, e->prod_mask));
        else MDK_CUDA(launch_rec_tc(ws.gi, fuse_x ? &fx : nullptr, e->layer[0].w_hh_tm, e->layer[0].b_hn_tc, ws.h0, 1, B, T,
                                    e->sm_count, s, nullptr, nullptr, e->prod_mask));
    } else {
        MDK_CUDA(launch_rec_fp32(ws.gi, e->layer[0].w_hh_t, e->layer[0].b_hn, (float *)ws.h0, B, T, s));
    }
    launches++;
    MDK_CUDA(cudaEventRecord(e->ev[3], s));
    if (tc) MDK_CUDA(launch_gemm_tc(ws.h0, e->layer[1].w_in_tc, e->layer[1].bias_gi_tc, ws.gi, tiled_rows(B, T), e->sm_count, s,
                                    e->prod_mask, ws.gemm_ctr));
    else MDK_CUDA(launch_gemm_fp32((const float *)ws.h0, e->layer[1].w_in_packed, e->layer[1].bias_gi, ws.gi, P, s));
    launches++;
    MDK_CUDA(cudaEventRecord(e->ev[4], s));
    if (tc) {
        if (pp && fuse_head) MDK_CUDA(launch_rec_pp(1, ws.gi, nullptr, e->layer[1].w_hh_tm, e->layer[1].b_hn_tc, nullptr, B, T, s,
                                                    e->lin_w_tc, ws.plog, e->prod_mask));
        else MDK_CUDA(launch_rec_tc(ws.gi, nullptr, e->layer[1].w_hh_tm, e->layer[1].b_hn_tc, ws.h1, 0, B, T, e->sm_count, s,
                                    fuse_head ? e->lin_w_tc : nullptr, fuse_head ? ws.plog : nullptr, e->prod_mask));
    } else {
        MDK_CUDA(launch_rec_fp32(ws.gi, e->layer[1].w_hh_t, e->layer[1].b_hn, ws.h1, B, T, s));
    }
    launches++;
    MDK_CUDA(cudaEventRecord(e->ev[5], s));
    if (fuse_head) MDK_CUDA(launch_head_plog(ws.plog, e->lin_b, B, T, probs_dev, logits_dev, labels_dev, s));
    else MDK_CUDA(launch_head(ws.h1, e->lin_w, e->lin_b, B, T, tc ? 1 : 0, probs_dev, logits_dev, labels_dev, s));
    launches++;
    MDK_CUDA(cudaEventRecord(e->ev[6], s));
    e->launches += launches;
    e->last.launches = launches;
    ws.last_fused_head = fuse_head;
    ws.last_B = B; ws.last_T = T; ws.last_precision = e->precision;
    e->last_ws = (int)(&ws - e->ws);
    return MDK_OK;
}

// Next lane for a forward of P positions, round robin within its size class; the lane's previous group must have left
// the device before its buffers are reused.
static int acquire_lane(mdk_engine *e, int64_t P, int *out) {
    int idx;
    if (P > mdk_engine::SMALL_POS) {
        idx = e->next_big;
        e->next_big = (e->next_big + 1) % mdk_engine::BIG_LANES;
    } else {
        idx = mdk_engine::BIG_LANES + e->next_small;
        e->next_small = (e->next_small + 1) % mdk_engine::SMALL_LANES;
    }
    mdk_lane &ln = e->lane[idx];
    if (ln.busy) {
        MDK_CUDA(cudaEventSynchronize(ln.ev_out));
        ln.busy = false;
    }
    *out = idx;
    return MDK_OK;
}

// Seal the open group: one forward over all of its windows, then the results back to each batch's host buffers.
static int launch_group(mdk_engine *e) {
    if (e->open_lane < 0) return MDK_OK;
    mdk_lane &ln = e->lane[e->open_lane];
    e->open_lane = -1;
    ln.open = false;
    if (ln.items.empty()) return MDK_OK;
    int rc;
    cudaStream_t s = ln.ws->stream;
    e->ev = e->evr[e->fwd_count % mdk_engine::EV_RING];
    e->fwd_count++;
    MDK_CUDA(cudaEventRecord(e->ev[0], s));
    MDK_CUDA(cudaEventRecord(ln.ev_in, e->copy_in));      // every feature copy of the group was queued on copy_in
    MDK_CUDA(cudaStreamWaitEvent(s, ln.ev_in, 0));
    if ((rc = run_forward(e, *ln.ws, ln.d_feats, ln.gB, ln.gT, ln.d_probs, ln.want_logits ? ln.d_logits : nullptr,
                          ln.want_labels ? ln.d_labels : nullptr)))
        return rc;
    MDK_CUDA(cudaEventRecord(ln.ev_done, s));
    MDK_CUDA(cudaEventRecord(e->ev[7], s));
    MDK_CUDA(cudaStreamWaitEvent(e->copy_out, ln.ev_done, 0));
    int64_t w0 = 0;
    for (const mdk_lane::Item &it : ln.items) {
        const size_t P = (size_t)it.B * ln.gT, off = (size_t)w0 * ln.gT;
        MDK_CUDA(cudaMemcpyAsync(it.probs, ln.d_probs + off * NCLS, P * NCLS * sizeof(float), cudaMemcpyDeviceToHost, e->copy_out));
        if (it.logits)
            MDK_CUDA(cudaMemcpyAsync(it.logits, ln.d_logits + off * NCLS, P * NCLS * sizeof(float), cudaMemcpyDeviceToHost, e->copy_out));
        if (it.labels) MDK_CUDA(cudaMemcpyAsync(it.labels, ln.d_labels + off, P, cudaMemcpyDeviceToHost, e->copy_out));
        w0 += it.B;
    }
    MDK_CUDA(cudaEventRecord(ln.ev_out, e->copy_out));
    ln.busy = true;
    return MDK_OK;
}

static float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    return ms;
}

static int check_shapes(const mdk_engine *e, const void *feats, int64_t B, int64_t T, const void *probs) {
    MDK_REQUIRE(e != nullptr, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(feats != nullptr && probs != nullptr, MDK_ERR_ARG, "forward: feats/probs must not be NULL");
    MDK_REQUIRE(B >= 1 && T >= 1, MDK_ERR_ARG, "forward: need B >= 1 and T >= 1");
    MDK_REQUIRE(B * T < (int64_t)1 << 40, MDK_ERR_ARG, "forward: B*T too large");
    return MDK_OK;
}

}  // namespace mdk

using namespace mdk;

extern "C" {

const char *mdk_plp_bases(void) { return "acgtACGTdD"; }
size_t mdk_featlen(void) { return 10; }
size_t mdk_fwd_del(void) { return 9; }
size_t mdk_rev_del(void) { return 8; }
const char *mdk_last_error(void) { return g_last_error.c_str(); }
const char *mdk_version(void) { return "medaka_b200 0.1.0 (sm_100a)"; }

int mdk_device_count(int *count) {
    MDK_REQUIRE(count, MDK_ERR_ARG, "count is NULL");
    cudaError_t err = cudaGetDeviceCount(count);
    if (err != cudaSuccess) { *count = 0; return cuda_fail(err, "cudaGetDeviceCount", __FILE__, __LINE__); }
    return MDK_OK;
}

int mdk_device_info(int device, int *sm_arch, int *sm_count, size_t *total_mem) {
    cudaDeviceProp prop;
    MDK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_arch) *sm_arch = prop.major * 10 + prop.minor;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return MDK_OK;
}

int mdk_host_alloc(size_t bytes, void **out) {
    MDK_REQUIRE(out, MDK_ERR_ARG, "out is NULL");
    MDK_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return MDK_OK;
}
int mdk_host_free(void *p) {
    if (p) MDK_CUDA(cudaFreeHost(p));
    return MDK_OK;
}
int mdk_dev_alloc(int device, size_t bytes, void **out) {
    MDK_REQUIRE(out, MDK_ERR_ARG, "out is NULL");
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaMalloc(out, bytes ? bytes : 1));
    return MDK_OK;
}
int mdk_dev_free(int device, void *p) {
    MDK_CUDA(cudaSetDevice(device));
    if (p) MDK_CUDA(cudaFree(p));
    return MDK_OK;
}
int mdk_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes) {
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return MDK_OK;
}
int mdk_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes) {
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return MDK_OK;
}
int mdk_dev_memset(int device, void *dst_dev, int value, size_t bytes) {
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaMemset(dst_dev, value, bytes));
    return MDK_OK;
}
int mdk_device_synchronize(int device) {
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaDeviceSynchronize());
    return MDK_OK;
}

int mdk_engine_create(int device, const mdk_model_desc *desc, mdk_engine **out) {
    MDK_REQUIRE(desc && out, MDK_ERR_ARG, "engine_create: NULL argument");
    MDK_REQUIRE(desc->gru_size == H, MDK_ERR_UNSUPPORTED, "engine_create: only gru_size == 128 is supported");
    MDK_REQUIRE(desc->n_layers == 2 && desc->bidirectional == 1, MDK_ERR_UNSUPPORTED,
                "engine_create: only the 2-layer bidirectional GRU is supported");
    MDK_REQUIRE(desc->num_features >= 1 && desc->num_features <= 1024, MDK_ERR_ARG, "engine_create: bad num_features");
    int ndev = 0;
    MDK_CUDA(cudaGetDeviceCount(&ndev));
    MDK_REQUIRE(device >= 0 && device < ndev, MDK_ERR_ARG, "engine_create: no such CUDA device");
    cudaDeviceProp prop;
    MDK_CUDA(cudaGetDeviceProperties(&prop, device));
    MDK_REQUIRE(prop.major == 10, MDK_ERR_UNSUPPORTED,
                "engine_create: this library is built for sm_100a (Blackwell B200) only");
    MDK_CUDA(cudaSetDevice(device));
    mdk_engine *e = new (std::nothrow) mdk_engine();
    MDK_REQUIRE(e, MDK_ERR_NOMEM, "engine_create: out of host memory");
    e->device = device;
    e->desc = *desc;
    e->sm_count = prop.multiProcessorCount;
    {
        const char *v = getenv("MDK_NO_FUSE_X");
        e->fuse_x = !(v && v[0] == '1');
        v = getenv("MDK_REC_MODE");          // A/B measurements: "one" / "pp"
        if (v && v[0] == 'o') e->rec_mode = MDK_REC_ONE_TILE;
        else if (v && v[0] == 'p') e->rec_mode = MDK_REC_PINGPONG;
        v = getenv("MDK_PRODUCTS");
        if (v && v[0] >= '1' && v[0] <= '7') e->prod_mask = ((uint32_t)(v[0] - '0') & 7u) | 1u;
    }
    for (auto &ws : e->ws) {
        cudaError_t err = cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking);
        if (err != cudaSuccess) { mdk_engine_destroy(e); return cuda_fail(err, "cudaStreamCreate", __FILE__, __LINE__); }
    }
    for (int i = 0; i < mdk_engine::N_LANES; ++i) {
        mdk_lane &ln = e->lane[i];
        ln.ws = i < mdk_engine::BIG_LANES ? &e->ws[i % mdk_engine::BIG_WS]
                                          : &e->ws[mdk_engine::BIG_WS + (i - mdk_engine::BIG_LANES)];
        cudaEventCreateWithFlags(&ln.ev_in, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.ev_out, cudaEventDisableTiming);
    }
    e->stream = e->ws[0].stream;
    for (auto &set : e->evr) for (auto &ev : set) cudaEventCreate(&ev);
    cudaStreamCreateWithFlags(&e->copy_in, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&e->copy_out, cudaStreamNonBlocking);
    for (auto &ev : e->ev_timer) cudaEventCreate(&ev);
    cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming);
    *out = e;
    return MDK_OK;
}

int mdk_engine_destroy(mdk_engine *e) {
    if (!e) return MDK_OK;
    cudaSetDevice(e->device);
    for (auto &ws : e->ws) if (ws.stream) cudaStreamSynchronize(ws.stream);
    if (e->copy_out) cudaStreamSynchronize(e->copy_out);
    for (int l = 0; l < 2; ++l) {
        LayerWeights &lw = e->layer[l];
        for (int d = 0; d < NDIR; ++d) { dev_free(lw.w_ih[d]); dev_free(lw.w_hh[d]); dev_free(lw.b_ih[d]); dev_free(lw.b_hh[d]); }
        dev_free(lw.w_in_packed); dev_free(lw.bias_gi); dev_free(lw.b_hn); dev_free(lw.w_hh_t);
        dev_free(lw.bias_gi_tc); dev_free(lw.b_hn_tc);
        dev_free(lw.w_hh_tm); dev_free(lw.w_x_tm); dev_free(lw.w_in_tc);
    }
    dev_free(e->lin_w); dev_free(e->lin_b); dev_free(e->lin_w_tc);
    for (auto &ws : e->ws) {
        dev_free(ws.gi); dev_free(ws.h1); dev_free(ws.plog);
        if (ws.gemm_ctr) { cudaFree(ws.gemm_ctr); ws.gemm_ctr = nullptr; }
        if (ws.h0) cudaFree(ws.h0);
        if (ws.stream) cudaStreamDestroy(ws.stream);
    }
    for (auto &ln : e->lane) {
        dev_free(ln.d_feats); dev_free(ln.d_probs); dev_free(ln.d_logits); dev_free(ln.d_labels);
        if (ln.ev_in) cudaEventDestroy(ln.ev_in);
        if (ln.ev_done) cudaEventDestroy(ln.ev_done);
        if (ln.ev_out) cudaEventDestroy(ln.ev_out);
    }
    if (e->copy_in) cudaStreamDestroy(e->copy_in);
    if (e->copy_out) cudaStreamDestroy(e->copy_out);
    for (auto &set : e->evr) for (auto &ev : set) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->ev_timer) if (ev) cudaEventDestroy(ev);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    delete e;
    cudaGetLastError();
    return MDK_OK;
}

// weights are read by every lane: quiesce all of them before the packed copies are rebuilt
static int quiesce(mdk_engine *e) {
    int rc = launch_group(e);
    if (rc) return rc;
    for (auto &ws : e->ws) MDK_CUDA(cudaStreamSynchronize(ws.stream));
    return MDK_OK;
}

int mdk_engine_load_gru(mdk_engine *e, int layer, int direction, const float *w_ih, const float *w_hh,
                        const float *b_ih, const float *b_hh) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(layer >= 0 && layer < 2 && direction >= 0 && direction < 2, MDK_ERR_ARG, "load_gru: bad layer/direction");
    MDK_REQUIRE(w_ih && w_hh && b_ih && b_hh, MDK_ERR_ARG, "load_gru: NULL weight pointer");
    MDK_CUDA(cudaSetDevice(e->device));
    { int rcq = quiesce(e); if (rcq) return rcq; }
    LayerWeights &lw = e->layer[layer];
    const int in = in_features(e, layer);
    int rc;
    if ((rc = upload(e->stream, &lw.w_ih[direction], w_ih, (size_t)G3 * in))) return rc;
    if ((rc = upload(e->stream, &lw.w_hh[direction], w_hh, (size_t)G3 * H))) return rc;
    if ((rc = upload(e->stream, &lw.b_ih[direction], b_ih, (size_t)G3))) return rc;
    if ((rc = upload(e->stream, &lw.b_hh[direction], b_hh, (size_t)G3))) return rc;
    lw.loaded[direction] = true;
    e->prepared = false;
    return MDK_OK;
}

int mdk_engine_load_linear(mdk_engine *e, const float *w, const float *b) {
    MDK_REQUIRE(e && w && b, MDK_ERR_ARG, "load_linear: NULL argument");
    MDK_CUDA(cudaSetDevice(e->device));
    { int rcq = quiesce(e); if (rcq) return rcq; }
    int rc;
    if ((rc = upload(e->stream, &e->lin_w, w, (size_t)NCLS * H2))) return rc;
    e->prepared = false;
    if ((rc = upload(e->stream, &e->lin_b, b, (size_t)NCLS))) return rc;
    e->lin_loaded = true;
    return MDK_OK;
}

int mdk_engine_set_precision(mdk_engine *e, int mode) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(mode == MDK_PREC_TC || mode == MDK_PREC_FP32, MDK_ERR_ARG, "set_precision: unknown mode");
    e->precision = mode;
    return MDK_OK;
}
int mdk_engine_get_precision(mdk_engine *e, int *mode) {
    MDK_REQUIRE(e && mode, MDK_ERR_ARG, "NULL argument");
    *mode = e->precision;
    return MDK_OK;
}

int mdk_engine_set_products(mdk_engine *e, int mask) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(mask >= 1 && mask <= 7 && (mask & 1), MDK_ERR_ARG,
                "set_products: mask is a bit set of {1: W_hi.x_hi (required), 2: W_hi.x_lo, 4: W_lo.x_hi}");
    e->prod_mask = (uint32_t)mask;
    return MDK_OK;
}

int mdk_engine_set_rec_mode(mdk_engine *e, int mode) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(mode == MDK_REC_AUTO || mode == MDK_REC_ONE_TILE || mode == MDK_REC_PINGPONG, MDK_ERR_ARG,
                "set_rec_mode: unknown mode");
    e->rec_mode = mode;
    return MDK_OK;
}

int mdk_engine_set_group_windows(mdk_engine *e, int64_t windows) {
    MDK_REQUIRE(e && windows >= 0, MDK_ERR_ARG, "set_group_windows: bad arguments");
    e->group_windows = windows;
    return MDK_OK;
}

// Reserve = size the big workspaces and the big lanes' staging for groups of up to B windows of T columns.  This is also
// what switches coalescing on: a group collects submitted batches only as far as its lane's staging buffers reach.
int mdk_engine_reserve(mdk_engine *e, int64_t B, int64_t T) {
    MDK_REQUIRE(e && B >= 1 && T >= 1, MDK_ERR_ARG, "reserve: bad arguments");
    MDK_CUDA(cudaSetDevice(e->device));
    int rc;
    if ((rc = launch_group(e))) return rc;
    const bool big = B * T > mdk_engine::SMALL_POS;
    const int l0 = big ? 0 : mdk_engine::BIG_LANES, l1 = big ? mdk_engine::BIG_LANES : mdk_engine::N_LANES;
    for (int i = l0; i < l1; ++i) {
        mdk_lane &ln = e->lane[i];
        if (ln.busy) { MDK_CUDA(cudaEventSynchronize(ln.ev_out)); ln.busy = false; }
        if ((rc = ensure_workspace(e, *ln.ws, B, T))) return rc;
        if ((rc = ensure_io(e, ln, B, T))) return rc;
    }
    return MDK_OK;
}

int mdk_engine_forward_dev(mdk_engine *e, const float *feats_dev, int64_t B, int64_t T, float *probs_dev,
                           float *logits_dev, uint8_t *labels_dev) {
    int rc;
    if ((rc = check_shapes(e, feats_dev, B, T, probs_dev))) return rc;
    MDK_CUDA(cudaSetDevice(e->device));
    if ((rc = launch_group(e))) return rc;
    // device buffers need no staging: take the next workspace of the size class (stream order keeps it safe)
    int wi;
    if (B * T > mdk_engine::SMALL_POS) {
        wi = e->next_big_ws;
        e->next_big_ws = (e->next_big_ws + 1) % mdk_engine::BIG_WS;
    } else {
        wi = mdk_engine::BIG_WS + e->next_small;
        e->next_small = (e->next_small + 1) % mdk_engine::SMALL_LANES;
    }
    mdk_ws &ws = e->ws[wi];
    e->ev = e->evr[e->fwd_count % mdk_engine::EV_RING];
    e->fwd_count++;
    MDK_CUDA(cudaEventRecord(e->ev[0], ws.stream));
    if ((rc = run_forward(e, ws, feats_dev, B, T, probs_dev, logits_dev, labels_dev))) return rc;
    MDK_CUDA(cudaEventRecord(e->ev[7], ws.stream));
    return MDK_OK;
}

// Submitted batches are packed into groups window by window: a batch that does not fit what is left of the open group
// is split (windows are independent, medaka/prediction.py:40-52 treats every row of a batch separately), so every
// group of a long run is exactly one wave however the caller sized its batches.  The ticket follows the batch's last
// piece; groups complete in submission order per lane and the pieces of one batch sit in consecutive groups.
int mdk_engine_submit(mdk_engine *e, const float *feats_host, int64_t B, int64_t T, float *probs_host,
                      float *logits_host, uint8_t *labels_host, int64_t *ticket) {
    int rc;
    if ((rc = check_shapes(e, feats_host, B, T, probs_host))) return rc;
    MDK_REQUIRE(ticket, MDK_ERR_ARG, "submit: ticket is NULL");
    MDK_CUDA(cudaSetDevice(e->device));
    const int64_t gmax = e->group_windows > 0 ? e->group_windows : mdk_engine_preferred_windows(e);
    const int64_t F = e->desc.num_features;
    if (e->open_lane >= 0 && e->lane[e->open_lane].gT != T && (rc = launch_group(e))) return rc;
    int64_t done = 0;
    while (done < B) {
        if (e->open_lane < 0) {
            int li;
            if ((rc = acquire_lane(e, B * T, &li))) return rc;      // size class of the BATCH: the tail piece of a split
                                                                    // batch stays with the big lanes
            mdk_lane &ln = e->lane[li];
            // the staging grows to this batch (at most one wave of it); a reserved lane is already larger and keeps
            // collecting further batches up to its size
            if ((rc = ensure_io(e, ln, std::min<int64_t>(B - done, gmax), T))) return rc;
            ln.items.clear();
            ln.gB = 0; ln.gT = T;
            ln.want_logits = false; ln.want_labels = false;
            ln.open = true;
            ln.group++;
            e->open_lane = li;
        }
        mdk_lane &ln = e->lane[e->open_lane];
        // windows the open group can still take: one wave, and what the lane's staging reaches
        int64_t room = std::min(gmax, std::min(ln.cap_io / T, ln.cap_feats / (T * F))) - ln.gB;
        if (ln.gB == 0 && room < 1) room = 1;        // (ensure_io above sized the staging for at least one window)
        if (room < 1) {
            if ((rc = launch_group(e))) return rc;
            continue;
        }
        const int64_t n = std::min(room, B - done);
        // copy-in stream: features H2D (asynchronous when feats_host is page-locked), behind the group's earlier pieces
        MDK_CUDA(cudaMemcpyAsync(ln.d_feats + (size_t)ln.gB * T * F, feats_host + (size_t)done * T * F,
                                 (size_t)n * T * F * sizeof(float), cudaMemcpyHostToDevice, e->copy_in));
        ln.items.push_back(mdk_lane::Item{feats_host + (size_t)done * T * F, probs_host + (size_t)done * T * NCLS,
                                          logits_host ? logits_host + (size_t)done * T * NCLS : nullptr,
                                          labels_host ? labels_host + (size_t)done * T : nullptr, n});
        ln.gB += n;
        ln.want_logits = ln.want_logits || logits_host != nullptr;
        ln.want_labels = ln.want_labels || labels_host != nullptr;
        done += n;
        if (done == B) {
            const int64_t tk = e->submit_count++;
            e->ticket_lane[tk % mdk_engine::TICKET_RING] = (int16_t)e->open_lane;
            e->ticket_group[tk % mdk_engine::TICKET_RING] = ln.group;
            *ticket = tk;
        }
        if (n == room && (rc = launch_group(e))) return rc;      // full: launch right away
    }
    return MDK_OK;
}

int mdk_engine_flush(mdk_engine *e) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_CUDA(cudaSetDevice(e->device));
    return launch_group(e);
}

int mdk_engine_wait(mdk_engine *e, int64_t ticket) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_REQUIRE(ticket >= 0 && ticket < e->submit_count, MDK_ERR_ARG, "wait: unknown ticket");
    if (ticket < e->submit_count - mdk_engine::TICKET_RING) return MDK_OK;   // its lane has been reused many times since
    MDK_CUDA(cudaSetDevice(e->device));
    const int li = e->ticket_lane[ticket % mdk_engine::TICKET_RING];
    const int64_t grp = e->ticket_group[ticket % mdk_engine::TICKET_RING];
    mdk_lane &ln = e->lane[li];
    if (ln.group != grp) return MDK_OK;        // a later group runs on the lane: this one was waited for when it was reused
    if (ln.open) {
        int rc = launch_group(e);              // still collecting: the caller wants the result now
        if (rc) return rc;
    }
    if (ln.busy) {
        MDK_CUDA(cudaEventSynchronize(ln.ev_out));
        ln.busy = false;
    }
    return MDK_OK;
}

int mdk_engine_forward(mdk_engine *e, const float *feats_host, int64_t B, int64_t T, float *probs_host,
                       float *logits_host, uint8_t *labels_host) {
    int64_t ticket = -1;
    int rc = mdk_engine_submit(e, feats_host, B, T, probs_host, logits_host, labels_host, &ticket);
    if (rc) return rc;
    return mdk_engine_wait(e, ticket);
}

int mdk_engine_sync(mdk_engine *e) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_CUDA(cudaSetDevice(e->device));
    int rc = launch_group(e);
    if (rc) return rc;
    MDK_CUDA(cudaStreamSynchronize(e->copy_in));
    for (auto &ws : e->ws) MDK_CUDA(cudaStreamSynchronize(ws.stream));
    MDK_CUDA(cudaStreamSynchronize(e->copy_out));
    for (auto &ln : e->lane) ln.busy = false;
    return MDK_OK;
}

static void stage_times(cudaEvent_t *ev, mdk_timings *t) {
    t->h2d_ms = ev_ms(ev[0], ev[1]);
    t->inproj0_ms = ev_ms(ev[1], ev[2]);
    t->rec0_ms = ev_ms(ev[2], ev[3]);
    t->inproj1_ms = ev_ms(ev[3], ev[4]);
    t->rec1_ms = ev_ms(ev[4], ev[5]);
    t->head_ms = ev_ms(ev[5], ev[6]);
    t->d2h_ms = ev_ms(ev[6], ev[7]);
    t->total_ms = ev_ms(ev[0], ev[7]);
}

int mdk_engine_last_timings(mdk_engine *e, mdk_timings *out) { return mdk_engine_mean_timings(e, 1, out); }

int mdk_engine_mean_timings(mdk_engine *e, int n_last, mdk_timings *out) {
    MDK_REQUIRE(e && out, MDK_ERR_ARG, "NULL argument");
    MDK_REQUIRE(n_last >= 1 && n_last <= mdk_engine::EV_RING, MDK_ERR_ARG, "mean_timings: n_last out of range");
    MDK_REQUIRE(e->fwd_count >= n_last, MDK_ERR_STATE, "mean_timings: fewer forwards recorded than requested");
    MDK_CUDA(cudaSetDevice(e->device));
    for (auto &ws : e->ws) MDK_CUDA(cudaStreamSynchronize(ws.stream));
    mdk_timings acc{};
    for (int i = 0; i < n_last; ++i) {
        mdk_timings t{};
        stage_times(e->evr[(e->fwd_count - 1 - i) % mdk_engine::EV_RING], &t);
        acc.h2d_ms += t.h2d_ms; acc.inproj0_ms += t.inproj0_ms; acc.rec0_ms += t.rec0_ms;
        acc.inproj1_ms += t.inproj1_ms; acc.rec1_ms += t.rec1_ms; acc.head_ms += t.head_ms;
        acc.d2h_ms += t.d2h_ms; acc.total_ms += t.total_ms;
    }
    const float inv = 1.0f / (float)n_last;
    acc.h2d_ms *= inv; acc.inproj0_ms *= inv; acc.rec0_ms *= inv; acc.inproj1_ms *= inv; acc.rec1_ms *= inv;
    acc.head_ms *= inv; acc.d2h_ms *= inv; acc.total_ms *= inv;
    acc.launches = e->last.launches;
    *out = acc;
    return MDK_OK;
}

// Diagnostics: completion times (ms after the timer's start event) of the eight stage events of the last n forwards,
// oldest first - the schedule the lanes actually ran (tools/diag.py --check timeline).
int mdk_debug_timeline(mdk_engine *e, int n_last, float *out) {
    MDK_REQUIRE(e && out, MDK_ERR_ARG, "NULL argument");
    MDK_REQUIRE(n_last >= 1 && n_last <= mdk_engine::EV_RING && e->fwd_count >= n_last, MDK_ERR_ARG, "timeline: bad n_last");
    MDK_CUDA(cudaSetDevice(e->device));
    for (auto &ws : e->ws) MDK_CUDA(cudaStreamSynchronize(ws.stream));
    for (int i = 0; i < n_last; ++i) {
        cudaEvent_t *ev = e->evr[(e->fwd_count - n_last + i) % mdk_engine::EV_RING];
        for (int k = 0; k < 8; ++k) MDK_CUDA(cudaEventElapsedTime(out + i * 8 + k, e->ev_timer[0], ev[k]));
    }
    return MDK_OK;
}

// The timed region spans every lane: the start event goes on lane 0 after all lanes have drained, the stop event on
// lane 0 after it has been made to wait for every other lane and for the copy-out stream.
int mdk_engine_timer_start(mdk_engine *e) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "engine is NULL");
    MDK_CUDA(cudaSetDevice(e->device));
    int rc = mdk_engine_sync(e);
    if (rc) return rc;
    MDK_CUDA(cudaEventRecord(e->ev_timer[0], e->stream));
    // work queued on the other lanes / copy streams after this point must not start before the start event
    for (int i = 1; i < mdk_engine::N_WS; ++i) MDK_CUDA(cudaStreamWaitEvent(e->ws[i].stream, e->ev_timer[0], 0));
    MDK_CUDA(cudaStreamWaitEvent(e->copy_in, e->ev_timer[0], 0));
    return MDK_OK;
}
int mdk_engine_timer_stop(mdk_engine *e, float *elapsed_ms) {
    MDK_REQUIRE(e && elapsed_ms, MDK_ERR_ARG, "NULL argument");
    MDK_CUDA(cudaSetDevice(e->device));
    int rc = launch_group(e);
    if (rc) return rc;
    for (int i = 1; i < mdk_engine::N_WS; ++i) {
        MDK_CUDA(cudaEventRecord(e->ev_join, e->ws[i].stream));
        MDK_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    }
    MDK_CUDA(cudaEventRecord(e->ev_join, e->copy_out));     // the end event must follow every copy-out still in flight
    MDK_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    MDK_CUDA(cudaEventRecord(e->ev_timer[1], e->stream));
    MDK_CUDA(cudaEventSynchronize(e->ev_timer[1]));
    MDK_CUDA(cudaEventElapsedTime(elapsed_ms, e->ev_timer[0], e->ev_timer[1]));
    return MDK_OK;
}

int mdk_engine_read_activation(mdk_engine *e, int which, float *out_host, int64_t n_floats) {
    MDK_REQUIRE(e && out_host, MDK_ERR_ARG, "NULL argument");
    MDK_REQUIRE(which == 0 || which == 1, MDK_ERR_ARG, "read_activation: which must be 0 or 1");
    mdk_ws &ln = e->ws[e->last_ws];
    const int64_t P = ln.last_B * ln.last_T;
    MDK_REQUIRE(P > 0 && n_floats == P * H2, MDK_ERR_ARG, "read_activation: size must be B*T*256 of the last forward");
    MDK_REQUIRE(!(which == 1 && ln.last_fused_head), MDK_ERR_STATE,
                "read_activation(1): the last forward fused the head into layer 1 (h1 never reached HBM); call "
                "mdk_engine_keep_activations(e, 1) before the forward");
    MDK_CUDA(cudaSetDevice(e->device));
    MDK_CUDA(cudaStreamSynchronize(ln.stream));
    if (ln.last_precision == MDK_PREC_FP32) {
        MDK_CUDA(cudaMemcpy(out_host, which == 1 ? (const void *)ln.h1 : (const void *)ln.h0,
                            (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        // tensor-core path: rows are tile-interleaved (and layer 0 is stored as fp16 hi/lo operand tiles)
        float *tmp = nullptr;
        MDK_CUDA(cudaMalloc(&tmp, (size_t)n_floats * sizeof(float)));
        cudaError_t err = which == 0 ? launch_unpack_h0(ln.h0, tmp, ln.last_B, ln.last_T, ln.stream)
                                     : launch_untile_rows(ln.h1, tmp, ln.last_B, ln.last_T, ln.stream);
        if (err == cudaSuccess) err = cudaStreamSynchronize(ln.stream);
        if (err == cudaSuccess) err = cudaMemcpy(out_host, tmp, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(tmp);
        e->launches++;
        if (err != cudaSuccess) return cuda_fail(err, "read_activation", __FILE__, __LINE__);
    }
    return MDK_OK;
}

int mdk_debug_pp_flags(int flags) {
    pp_set_debug((uint32_t)flags);
    return MDK_OK;
}

int mdk_debug_read_plog(mdk_engine *e, float *out_host, int64_t n_floats) {
    MDK_REQUIRE(e && out_host, MDK_ERR_ARG, "NULL argument");
    mdk_ws &ln = e->ws[e->last_ws];
    MDK_REQUIRE(ln.last_fused_head, MDK_ERR_STATE, "read_plog: the last forward did not run the fused head");
    const int64_t tiles = (ln.last_B + WT - 1) / WT;
    MDK_REQUIRE(n_floats == NDIR * tiles * ln.last_T * PLOG_TS_FLOATS, MDK_ERR_ARG, "read_plog: size must be 2*tiles*T*80");
    MDK_CUDA(cudaSetDevice(e->device));
    MDK_CUDA(cudaStreamSynchronize(ln.stream));
    MDK_CUDA(cudaMemcpy(out_host, ln.plog, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost));
    return MDK_OK;
}

int64_t mdk_engine_launch_count(mdk_engine *e) { return e ? e->launches : 0; }

int mdk_engine_keep_activations(mdk_engine *e, int keep) {
    MDK_REQUIRE(e, MDK_ERR_ARG, "NULL engine");
    e->keep_act = keep != 0;
    return MDK_OK;
}

int64_t mdk_engine_preferred_windows(mdk_engine *e) {
    const int sms = e ? e->sm_count : 148;
    return (int64_t)mdk::WT * (sms / mdk::NDIR);
}

// ---------------------------------------------------------------------------- featuriser seam
int mdk_normalise_counts_dev(int device, const uint64_t *counts_dev, const int64_t *major_dev,
                             const int64_t *minor_dev, int64_t n, int32_t num_dtypes, int32_t mode,
                             int32_t sym_indels, float *feats_out_dev, int64_t *depth_out_dev) {
    MDK_REQUIRE(n >= 0, MDK_ERR_ARG, "normalise_counts: n < 0");
    MDK_REQUIRE(num_dtypes >= 1 && num_dtypes <= 4, MDK_ERR_UNSUPPORTED, "normalise_counts: 1..4 dtypes supported");
    MDK_REQUIRE(mode >= MDK_NORM_TOTAL && mode <= MDK_NORM_NONE, MDK_ERR_ARG, "normalise_counts: unknown mode");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(counts_dev && major_dev && minor_dev && feats_out_dev, MDK_ERR_ARG, "normalise_counts: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(launch_normalise(counts_dev, major_dev, minor_dev, n, num_dtypes, mode, sym_indels, feats_out_dev,
                              depth_out_dev, 0));
    return MDK_OK;
}

int mdk_normalise_counts(int device, const uint64_t *counts, const int64_t *major, const int64_t *minor, int64_t n,
                         int32_t num_dtypes, int32_t mode, int32_t sym_indels, float *feats_out,
                         int64_t *depth_out) {
    MDK_REQUIRE(n >= 0, MDK_ERR_ARG, "normalise_counts: n < 0");
    MDK_REQUIRE(num_dtypes >= 1 && num_dtypes <= 4, MDK_ERR_UNSUPPORTED, "normalise_counts: 1..4 dtypes supported");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(counts && major && minor && feats_out, MDK_ERR_ARG, "normalise_counts: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    const size_t F = 10 * (size_t)num_dtypes;
    uint8_t *buf = nullptr;
    const size_t b_counts = (size_t)n * F * 8, b_pos = (size_t)n * 8, b_feats = (size_t)n * F * 4;
    MDK_CUDA(plp_scratch(b_counts + 3 * b_pos + b_feats + 64, &buf, 1));   // cached per host thread (the loader threads call this per region)
    uint64_t *d_counts = reinterpret_cast<uint64_t *>(buf);
    int64_t *d_major = reinterpret_cast<int64_t *>(buf + b_counts);
    int64_t *d_minor = d_major + n;
    int64_t *d_depth = d_minor + n;
    float *d_feats = reinterpret_cast<float *>(buf + b_counts + 3 * b_pos);
    cudaError_t err = cudaMemcpy(d_counts, counts, b_counts, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_major, major, b_pos, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_minor, minor, b_pos, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = launch_normalise(d_counts, d_major, d_minor, n, num_dtypes, mode, sym_indels, d_feats, d_depth, 0);
    if (err == cudaSuccess) err = cudaMemcpy(feats_out, d_feats, b_feats, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && depth_out) err = cudaMemcpy(depth_out, d_depth, b_pos, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return cuda_fail(err, "normalise_counts", __FILE__, __LINE__);
    return MDK_OK;
}

int mdk_pileup_counts(int device, int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                      const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, const uint8_t *seq,
                      const int64_t *seq_off, int32_t start, int32_t end, int32_t num_dtypes, int32_t min_mapq,
                      int64_t max_cols, uint64_t *counts_out, int64_t *major_out, int64_t *minor_out,
                      int64_t *n_cols_out) {
    MDK_REQUIRE(n_cols_out, MDK_ERR_ARG, "pileup_counts: n_cols_out is NULL");
    *n_cols_out = 0;
    MDK_REQUIRE(n_rec >= 0 && end >= start && max_cols >= 0, MDK_ERR_ARG, "pileup_counts: bad sizes");
    MDK_REQUIRE(num_dtypes >= 1 && num_dtypes <= 4, MDK_ERR_UNSUPPORTED, "pileup_counts: 1..4 dtypes supported");
    if (n_rec == 0 || end == start) return MDK_OK;
    MDK_REQUIRE(pos && flag && mapq && dtype && cigar && cigar_off && seq && seq_off, MDK_ERR_ARG,
                "pileup_counts: NULL record array");
    MDK_REQUIRE(max_cols == 0 || (counts_out && major_out && minor_out), MDK_ERR_ARG, "pileup_counts: NULL output");
    MDK_CUDA(cudaSetDevice(device));
    const int64_t n_ops = cigar_off[n_rec], n_seq = seq_off[n_rec];
    const int F = 10 * num_dtypes;
    // one device blob: records in, columns out
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
    const size_t o_pos = take((size_t)n_rec * 4), o_flag = take((size_t)n_rec * 2), o_mapq = take((size_t)n_rec),
                 o_dt = take((size_t)n_rec), o_cig = take((size_t)n_ops * 4), o_coff = take((size_t)(n_rec + 1) * 8),
                 o_seq = take((size_t)n_seq), o_soff = take((size_t)(n_rec + 1) * 8),
                 o_cnt = take((size_t)max_cols * F * 8), o_maj = take((size_t)max_cols * 8),
                 o_min = take((size_t)max_cols * 8);
    uint8_t *buf = nullptr;
    MDK_CUDA(plp_scratch(off + 16, &buf, 1));     // cached per host thread: no cudaMalloc / cudaFree per region
    cudaStream_t s = 0;
    cudaError_t err = cudaMemcpy(buf + o_pos, pos, (size_t)n_rec * 4, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(buf + o_flag, flag, (size_t)n_rec * 2, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(buf + o_mapq, mapq, (size_t)n_rec, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(buf + o_dt, dtype, (size_t)n_rec, cudaMemcpyHostToDevice);
    if (err == cudaSuccess && n_ops) err = cudaMemcpy(buf + o_cig, cigar, (size_t)n_ops * 4, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(buf + o_coff, cigar_off, (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice);
    if (err == cudaSuccess && n_seq) err = cudaMemcpy(buf + o_seq, seq, (size_t)n_seq, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(buf + o_soff, seq_off, (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice);
    int rc = MDK_OK;
    if (err == cudaSuccess) {
        rc = pileup_counts_dev(n_rec, (const int32_t *)(buf + o_pos), (const uint16_t *)(buf + o_flag), buf + o_mapq,
                               buf + o_dt, (const uint32_t *)(buf + o_cig), (const int64_t *)(buf + o_coff), n_ops,
                               buf + o_seq, (const int64_t *)(buf + o_soff), start, end, num_dtypes, min_mapq, max_cols,
                               (uint64_t *)(buf + o_cnt), (int64_t *)(buf + o_maj), (int64_t *)(buf + o_min),
                               n_cols_out, s);
    }
    if (rc == MDK_OK && err == cudaSuccess && *n_cols_out > max_cols) {
        set_error("pileup_counts: output buffers too small (see *n_cols_out)");
        rc = MDK_ERR_NOMEM;
    } else if (rc == MDK_OK && err == cudaSuccess && *n_cols_out > 0) {
        const int64_t n = *n_cols_out;
        err = cudaMemcpy(counts_out, buf + o_cnt, (size_t)n * F * 8, cudaMemcpyDeviceToHost);
        if (err == cudaSuccess) err = cudaMemcpy(major_out, buf + o_maj, (size_t)n * 8, cudaMemcpyDeviceToHost);
        if (err == cudaSuccess) err = cudaMemcpy(minor_out, buf + o_min, (size_t)n * 8, cudaMemcpyDeviceToHost);
    }
    if (err != cudaSuccess) return cuda_fail(err, "pileup_counts", __FILE__, __LINE__);
    return rc;
}

// Fused featuriser: records -> counts -> normalised features without the counts ever leaving the device (SURVEY.md 8f row
// f3: "a1 -> a3 fused").  Same arguments as mdk_pileup_counts plus the normalisation switches of mdk_normalise_counts;
// copies out 64 B per column (F = 10: features 40, depth 8, positions 16) instead of 96 B out, 96 B back in and 48 B out.
int mdk_pileup_features(int device, int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                        const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, const uint8_t *seq,
                        const int64_t *seq_off, int32_t start, int32_t end, int32_t num_dtypes, int32_t min_mapq,
                        int32_t mode, int32_t sym_indels, int64_t max_cols, float *feats_out, int64_t *depth_out,
                        int64_t *major_out, int64_t *minor_out, int64_t *n_cols_out) {
    MDK_REQUIRE(n_cols_out, MDK_ERR_ARG, "pileup_features: n_cols_out is NULL");
    *n_cols_out = 0;
    MDK_REQUIRE(n_rec >= 0 && end >= start && max_cols >= 0, MDK_ERR_ARG, "pileup_features: bad sizes");
    MDK_REQUIRE(num_dtypes >= 1 && num_dtypes <= 4, MDK_ERR_UNSUPPORTED, "pileup_features: 1..4 dtypes supported");
    MDK_REQUIRE(mode >= MDK_NORM_TOTAL && mode <= MDK_NORM_NONE, MDK_ERR_ARG, "pileup_features: unknown mode");
    if (n_rec == 0 || end == start) return MDK_OK;
    MDK_REQUIRE(pos && flag && mapq && dtype && cigar && cigar_off && seq && seq_off, MDK_ERR_ARG,
                "pileup_features: NULL record array");
    MDK_REQUIRE(max_cols == 0 || (feats_out && major_out && minor_out), MDK_ERR_ARG, "pileup_features: NULL output");
    MDK_CUDA(cudaSetDevice(device));
    const int64_t n_ops = cigar_off[n_rec], n_seq = seq_off[n_rec];
    const int F = 10 * num_dtypes;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
    const size_t o_pos = take((size_t)n_rec * 4), o_flag = take((size_t)n_rec * 2), o_mapq = take((size_t)n_rec),
                 o_dt = take((size_t)n_rec), o_cig = take((size_t)n_ops * 4), o_coff = take((size_t)(n_rec + 1) * 8),
                 o_seq = take((size_t)n_seq), o_soff = take((size_t)(n_rec + 1) * 8),
                 o_cnt = take((size_t)max_cols * F * 8), o_maj = take((size_t)max_cols * 8),
                 o_min = take((size_t)max_cols * 8), o_feat = take((size_t)max_cols * F * 4),
                 o_dep = take((size_t)max_cols * 8);
    uint8_t *buf = nullptr;
    MDK_CUDA(plp_scratch(off + 16, &buf, 1));
    cudaError_t err = cudaSuccess;
    auto up = [&](size_t o, const void *src, size_t bytes) {
        if (err == cudaSuccess && bytes) err = cudaMemcpy(buf + o, src, bytes, cudaMemcpyHostToDevice);
    };
    up(o_pos, pos, (size_t)n_rec * 4); up(o_flag, flag, (size_t)n_rec * 2); up(o_mapq, mapq, (size_t)n_rec);
    up(o_dt, dtype, (size_t)n_rec); up(o_cig, cigar, (size_t)n_ops * 4); up(o_coff, cigar_off, (size_t)(n_rec + 1) * 8);
    up(o_seq, seq, (size_t)n_seq); up(o_soff, seq_off, (size_t)(n_rec + 1) * 8);
    if (err != cudaSuccess) return cuda_fail(err, "pileup_features (copy in)", __FILE__, __LINE__);
    int rc = pileup_counts_dev(n_rec, (const int32_t *)(buf + o_pos), (const uint16_t *)(buf + o_flag), buf + o_mapq,
                               buf + o_dt, (const uint32_t *)(buf + o_cig), (const int64_t *)(buf + o_coff), n_ops,
                               buf + o_seq, (const int64_t *)(buf + o_soff), start, end, num_dtypes, min_mapq, max_cols,
                               (uint64_t *)(buf + o_cnt), (int64_t *)(buf + o_maj), (int64_t *)(buf + o_min), n_cols_out, 0);
    if (rc) return rc;
    const int64_t n = *n_cols_out;
    if (n > max_cols) {
        set_error("pileup_features: output buffers too small (see *n_cols_out)");
        return MDK_ERR_NOMEM;
    }
    if (n == 0) return MDK_OK;
    MDK_CUDA(launch_normalise((const uint64_t *)(buf + o_cnt), (const int64_t *)(buf + o_maj), (const int64_t *)(buf + o_min), n,
                              num_dtypes, mode, sym_indels, (float *)(buf + o_feat), (int64_t *)(buf + o_dep), 0));
    err = cudaMemcpy(feats_out, buf + o_feat, (size_t)n * F * 4, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && depth_out) err = cudaMemcpy(depth_out, buf + o_dep, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess) err = cudaMemcpy(major_out, buf + o_maj, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess) err = cudaMemcpy(minor_out, buf + o_min, (size_t)n * 8, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return cuda_fail(err, "pileup_features (copy out)", __FILE__, __LINE__);
    return MDK_OK;
}

// ---------------------------------------------------------------------------- decode seam
int mdk_decode_consensus_dev(int device, const float *probs_dev, int64_t n, uint8_t *labels_out_dev,
                             uint8_t *quals_out_dev) {
    MDK_REQUIRE(n >= 0, MDK_ERR_ARG, "decode_consensus: n < 0");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(probs_dev && labels_out_dev, MDK_ERR_ARG, "decode_consensus: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(launch_decode(probs_dev, n, labels_out_dev, quals_out_dev, 0));
    return MDK_OK;
}

int mdk_decode_consensus(int device, const float *probs, int64_t n, uint8_t *labels_out, uint8_t *quals_out) {
    MDK_REQUIRE(n >= 0, MDK_ERR_ARG, "decode_consensus: n < 0");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(probs && labels_out, MDK_ERR_ARG, "decode_consensus: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    uint8_t *buf = nullptr;
    const size_t b_probs = (size_t)n * NCLS * 4;
    MDK_CUDA(cudaMalloc(&buf, b_probs + 2 * (size_t)n));
    float *d_probs = reinterpret_cast<float *>(buf);
    uint8_t *d_labels = buf + b_probs, *d_quals = d_labels + n;
    cudaError_t err = cudaMemcpy(d_probs, probs, b_probs, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = launch_decode(d_probs, n, d_labels, quals_out ? d_quals : nullptr, 0);
    if (err == cudaSuccess) err = cudaMemcpy(labels_out, d_labels, (size_t)n, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && quals_out) err = cudaMemcpy(quals_out, d_quals, (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (err != cudaSuccess) return cuda_fail(err, "decode_consensus", __FILE__, __LINE__);
    return MDK_OK;
}

int mdk_decode_consensus_f64(int device, const double *probs, int64_t n, uint8_t *labels_out, uint8_t *quals_out) {
    MDK_REQUIRE(n >= 0, MDK_ERR_ARG, "decode_consensus: n < 0");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(probs && labels_out, MDK_ERR_ARG, "decode_consensus: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    uint8_t *buf = nullptr;
    const size_t b_probs = (size_t)n * NCLS * 8;
    MDK_CUDA(cudaMalloc(&buf, b_probs + 2 * (size_t)n));
    double *d_probs = reinterpret_cast<double *>(buf);
    uint8_t *d_labels = buf + b_probs, *d_quals = d_labels + n;
    cudaError_t err = cudaMemcpy(d_probs, probs, b_probs, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = launch_decode_f64(d_probs, n, d_labels, quals_out ? d_quals : nullptr, 0);
    if (err == cudaSuccess) err = cudaMemcpy(labels_out, d_labels, (size_t)n, cudaMemcpyDeviceToHost);
    if (err == cudaSuccess && quals_out) err = cudaMemcpy(quals_out, d_quals, (size_t)n, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (err != cudaSuccess) return cuda_fail(err, "decode_consensus_f64", __FILE__, __LINE__);
    return MDK_OK;
}

int mdk_variant_columns(int device, const int64_t *minor, const uint8_t *reference, const uint8_t *prediction,
                        uint8_t *out, int64_t len) {
    MDK_REQUIRE(len >= 0, MDK_ERR_ARG, "variant_columns: len < 0");
    if (len == 0) return MDK_OK;
    MDK_REQUIRE(minor && reference && prediction && out, MDK_ERR_ARG, "variant_columns: NULL pointer");
    MDK_CUDA(cudaSetDevice(device));
    uint8_t *buf = nullptr;
    const size_t n = (size_t)len;
    MDK_CUDA(cudaMalloc(&buf, n * 8 + 3 * n + 64));
    int64_t *d_minor = reinterpret_cast<int64_t *>(buf);
    uint8_t *d_ref = buf + n * 8, *d_pred = d_ref + n, *d_out = d_pred + n;
    cudaError_t err = cudaMemcpy(d_minor, minor, n * 8, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_ref, reference, n, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(d_pred, prediction, n, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = launch_variant_columns(d_minor, d_ref, d_pred, len, d_out, 0);
    if (err == cudaSuccess) err = cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost);
    cudaFree(buf);
    if (err != cudaSuccess) return cuda_fail(err, "variant_columns", __FILE__, __LINE__);
    return MDK_OK;
}

int mdk_selftest_umma(int device, const float *A, const float *B, float *D, int N, int K, int variant) {
    MDK_REQUIRE(A && B && D, MDK_ERR_ARG, "selftest_umma: NULL pointer");
    return selftest_umma(device, A, B, D, N, K, variant);
}

int mdk_debug_rec_trace(int device, int enable, uint64_t *out) {
    MDK_CUDA(cudaSetDevice(device));
    MDK_CUDA(cudaDeviceSynchronize());
    MDK_CUDA(rec_trace_control(enable, reinterpret_cast<unsigned long long *>(out)));
    return MDK_OK;
}

}  // extern "C"
