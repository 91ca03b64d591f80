// medaka_b200: variant decoding on the device.
//
// What medaka does per joined sample (medaka/labels.py:889-1014 `decode_variants`): argmax-decode the [n,5] label
// probabilities keeping gaps, lay the draft out with '*' on insertion columns, mark the variant columns
// (src/medaka_rnn_variants.c:28-55), cut them into runs and give every run the log-likelihood-ratio quality
//     sum_i phred(1 - p[i][pred_i]) - sum_i phred(1 - p[i][ref_i])        (labels.py:957-975, 387-401)
// summed left to right in the precision of the probabilities (float32 in production).  String building and VCF
// normalisation (Variant.trim, vcf.py:338-402) are O(#variants) and stay on the host.
//
//   vd_decode_kernel   column -> argmax label, mismatch flag, phred of the predicted and of the reference class
//   vd_group_kernel    variant_columns: an insertion column is variant when any column of its reference position is
//   vd_starts_kernel   run starts (variant column whose left neighbour is not) counted per block
//   stitch-style scan  exclusive scan of the block counts (single block)
//   vd_runs_kernel     the k-th run start walks its run: length and the two left-to-right float32 sums
//
// HBM-bound byte work: 29 B read + 10 B written per column in the first kernel, ~12 B per column in the others.
#include "common.cuh"

namespace mdk {

namespace {

constexpr int VD_THREADS = 256;

__device__ __forceinline__ float phred_f32(float p_class) {
    // labels.py:387-401 on float32: err = clip(1 - p, 1e-7, 1); q = min(-10 log10(err), 70)
    const float err = fminf(fmaxf(1.0f - p_class, 1e-7f), 1.0f);
    const float l = __double2float_rn(log10((double)err));   // correctly rounded float32 log10
    return fminf(-10.0f * l, 70.0f);
}

// ref_code: 0..4 = '*ACGT' (labels.py:342); 5 = 'N' (compared as a symbol of its own, scored as '*': labels.py:949-952);
// >= 6 = any other draft symbol (never equal to a call; scored as '*' - the host refuses it unless the run is skipped)
__global__ void __launch_bounds__(VD_THREADS) vd_decode_kernel(const float *__restrict__ probs,
                                                               const uint8_t *__restrict__ ref_code, int64_t n,
                                                               uint8_t *__restrict__ pred, uint8_t *__restrict__ mism,
                                                               float *__restrict__ pred_q, float *__restrict__ ref_q) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[NCLS];
#pragma unroll
    for (int c = 0; c < NCLS; ++c) p[c] = __ldcs(probs + i * NCLS + c);
    float best = p[0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < NCLS; ++c)
        if (p[c] > best) { best = p[c]; arg = c; }          // first maximum wins (np.argmax, labels.py:1063)
    const int r = ref_code[i];
    const int rq = r < NCLS ? r : 0;
    float pr = p[0];
#pragma unroll
    for (int c = 1; c < NCLS; ++c) pr = (rq == c) ? p[c] : pr;
    pred[i] = (uint8_t)arg;
    mism[i] = (uint8_t)(r != arg);
    pred_q[i] = phred_f32(best);
    ref_q[i] = phred_f32(pr);
}

// src/medaka_rnn_variants.c:28-55 on mismatch flags: a major column is variant when it mismatches; the insertion
// columns that follow it are variant when ANY column of the group (the major or one of its minors) mismatches.
__global__ void __launch_bounds__(VD_THREADS) vd_group_kernel(const int64_t *__restrict__ minor,
                                                              const uint8_t *__restrict__ mism, int64_t n,
                                                              uint8_t *__restrict__ is_var) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool any = mism[i] != 0;
    if (i != 0 && minor[i] != 0) {
        for (int64_t j = i - 1; j >= 0 && !any; --j) {      // back to (and including) the group's major column
            any = mism[j] != 0;
            if (j == 0 || minor[j] == 0) break;
        }
        for (int64_t j = i + 1; j < n && !any && minor[j] != 0; ++j) any = mism[j] != 0;
    }
    is_var[i] = any;
}

__global__ void __launch_bounds__(VD_THREADS) vd_starts_kernel(const uint8_t *__restrict__ is_var, int64_t n,
                                                               uint32_t *__restrict__ block_count) {
    __shared__ uint32_t warp_cnt[VD_THREADS / 32];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    uint32_t start = 0;
    if (i < n) start = is_var[i] && (i == 0 || !is_var[i - 1]);
    const uint32_t ballot = __ballot_sync(0xffffffffu, start);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(ballot);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < VD_THREADS / 32; ++w) t += warp_cnt[w];
        block_count[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts (same single-block walk as stitch_scan_kernel); base[n_blocks] = total
__global__ void __launch_bounds__(1024) vd_scan_kernel(const uint32_t *__restrict__ block_count, int64_t n_blocks,
                                                       int64_t *__restrict__ block_base) {
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t s = 0; s < n_blocks; s += 1024) {
        const int64_t i = s + threadIdx.x;
        const int64_t v = i < n_blocks ? (int64_t)block_count[i] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            warp_sum[lane] = w;
        }
        __syncthreads();
        const int64_t before = carry + (warp ? warp_sum[warp - 1] : 0) + (x - v);
        if (i < n_blocks) block_base[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_base[n_blocks] = carry;
}

__global__ void __launch_bounds__(VD_THREADS) vd_runs_kernel(const uint8_t *__restrict__ is_var,
                                                             const float *__restrict__ pred_q,
                                                             const float *__restrict__ ref_q, int64_t n,
                                                             const int64_t *__restrict__ block_base, int64_t max_runs,
                                                             int64_t *__restrict__ run_start, int64_t *__restrict__ run_len,
                                                             float *__restrict__ run_pred_q, float *__restrict__ run_ref_q) {
    __shared__ uint32_t warp_cnt[VD_THREADS / 32];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    uint32_t start = 0;
    if (i < n) start = is_var[i] && (i == 0 || !is_var[i - 1]);
    const uint32_t ballot = __ballot_sync(0xffffffffu, start);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    if (!start) return;
    uint32_t before = __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    const int64_t k = block_base[blockIdx.x] + before;
    if (k >= max_runs) return;
    // Python's sum(): start from int 0, add left to right - float32 + float32 in float32
    float sp = 0.0f, sr = 0.0f;
    int64_t j = i;
    for (; j < n && is_var[j]; ++j) {
        sp += pred_q[j];
        sr += ref_q[j];
    }
    run_start[k] = i;
    run_len[k] = j - i;
    run_pred_q[k] = sp;
    run_ref_q[k] = sr;
}

}  // namespace

}  // namespace mdk

using namespace mdk;

extern "C" {

int mdk_decode_variants(int device, const float *probs, const int64_t *minor, const uint8_t *ref_code, int64_t n,
                        uint8_t *pred_out, uint8_t *is_var_out, float *pred_q_out, float *ref_q_out, int64_t max_runs,
                        int64_t *run_start, int64_t *run_len, float *run_pred_q, float *run_ref_q,
                        int64_t *n_runs_out) {
    MDK_REQUIRE(n_runs_out, MDK_ERR_ARG, "decode_variants: n_runs_out is NULL");
    *n_runs_out = 0;
    MDK_REQUIRE(n >= 0 && max_runs >= 0, MDK_ERR_ARG, "decode_variants: negative size");
    if (n == 0) return MDK_OK;
    MDK_REQUIRE(probs && minor && ref_code && pred_out && is_var_out, MDK_ERR_ARG, "decode_variants: NULL pointer");
    MDK_REQUIRE(max_runs == 0 || (run_start && run_len && run_pred_q && run_ref_q), MDK_ERR_ARG,
                "decode_variants: NULL run output");
    MDK_REQUIRE(minor[0] == 0, MDK_ERR_ARG,
                "decode_variants: the first position of a sample must not be an insertion (labels.py:909-911)");
    MDK_CUDA(cudaSetDevice(device));
    const int64_t n_blocks = (n + VD_THREADS - 1) / VD_THREADS;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += (bytes + 15) / 16 * 16; return o; };
    const size_t o_probs = take((size_t)n * NCLS * 4), o_minor = take((size_t)n * 8), o_ref = take((size_t)n),
                 o_pred = take((size_t)n), o_mism = take((size_t)n), o_var = take((size_t)n),
                 o_pq = take((size_t)n * 4), o_rq = take((size_t)n * 4), o_cnt = take((size_t)n_blocks * 4),
                 o_base = take((size_t)(n_blocks + 1) * 8), o_rs = take((size_t)max_runs * 8),
                 o_rl = take((size_t)max_runs * 8), o_rp = take((size_t)max_runs * 4), o_rr = take((size_t)max_runs * 4);
    uint8_t *buf = nullptr;
    MDK_CUDA(plp_scratch(off + 16, &buf, 4));     // cached per host thread: no cudaMalloc / cudaFree per sample
    cudaStream_t s = 0;
    cudaError_t err = cudaMemcpyAsync(buf + o_probs, probs, (size_t)n * NCLS * 4, cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(buf + o_minor, minor, (size_t)n * 8, cudaMemcpyHostToDevice, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(buf + o_ref, ref_code, (size_t)n, cudaMemcpyHostToDevice, s);
    int64_t total = 0;
    if (err == cudaSuccess) {
        const unsigned g = (unsigned)n_blocks;
        vd_decode_kernel<<<g, VD_THREADS, 0, s>>>((const float *)(buf + o_probs), buf + o_ref, n, buf + o_pred, buf + o_mism,
                                                  (float *)(buf + o_pq), (float *)(buf + o_rq));
        vd_group_kernel<<<g, VD_THREADS, 0, s>>>((const int64_t *)(buf + o_minor), buf + o_mism, n, buf + o_var);
        vd_starts_kernel<<<g, VD_THREADS, 0, s>>>(buf + o_var, n, (uint32_t *)(buf + o_cnt));
        vd_scan_kernel<<<1, 1024, 0, s>>>((const uint32_t *)(buf + o_cnt), n_blocks, (int64_t *)(buf + o_base));
        vd_runs_kernel<<<g, VD_THREADS, 0, s>>>(buf + o_var, (const float *)(buf + o_pq), (const float *)(buf + o_rq), n,
                                                (const int64_t *)(buf + o_base), max_runs, (int64_t *)(buf + o_rs),
                                                (int64_t *)(buf + o_rl), (float *)(buf + o_rp), (float *)(buf + o_rr));
        err = cudaGetLastError();
    }
    if (err == cudaSuccess)
        err = cudaMemcpyAsync(&total, buf + o_base + (size_t)n_blocks * 8, 8, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(pred_out, buf + o_pred, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaMemcpyAsync(is_var_out, buf + o_var, (size_t)n, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess && pred_q_out)
        err = cudaMemcpyAsync(pred_q_out, buf + o_pq, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess && ref_q_out)
        err = cudaMemcpyAsync(ref_q_out, buf + o_rq, (size_t)n * 4, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    int rc = MDK_OK;
    if (err == cudaSuccess) {
        *n_runs_out = total;
        if (total > max_runs) {
            set_error("decode_variants: run buffers too small (see *n_runs_out)");
            rc = MDK_ERR_NOMEM;
        } else if (total > 0) {
            err = cudaMemcpy(run_start, buf + o_rs, (size_t)total * 8, cudaMemcpyDeviceToHost);
            if (err == cudaSuccess) err = cudaMemcpy(run_len, buf + o_rl, (size_t)total * 8, cudaMemcpyDeviceToHost);
            if (err == cudaSuccess) err = cudaMemcpy(run_pred_q, buf + o_rp, (size_t)total * 4, cudaMemcpyDeviceToHost);
            if (err == cudaSuccess) err = cudaMemcpy(run_ref_q, buf + o_rr, (size_t)total * 4, cudaMemcpyDeviceToHost);
        }
    }
    
    if (err != cudaSuccess) return cuda_fail(err, "decode_variants", __FILE__, __LINE__);
    return rc;
}

}  // extern "C"
