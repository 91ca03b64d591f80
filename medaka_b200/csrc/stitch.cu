// medaka_b200: consensus stitching on the device.
//
// What medaka does per region (medaka/stitch.py:33-85): for every trimmed sample, argmax-decode the [n,5] label
// probabilities, compute the phred quality of the winning class, drop the gap calls and append the survivors to the
// contig.  Here the host plans the kept row ranges (medaka_b200/stitch.py), all ranges of a call are laid out back
// to back in one device array, and three launches produce the final FASTA/FASTQ bytes:
//
//   stitch_decode_kernel    row -> (ASCII base | 0 for gap, quality char); per-block survivor count
//   stitch_scan_kernel      exclusive scan of the block counts (single block; <= a few 10^4 entries)
//   stitch_scatter_kernel   stable compaction: survivors move to block_base + rank-in-block; the first row of every
//                           range records where that range's output starts
//
// HBM-bound byte work: 20 B read per row + 2 B scratch written, 2 B scratch read + <= 2 B written.
#include "common.cuh"

#include <vector>

namespace mdk {

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_ROWS_PER_THREAD = 4;
constexpr int ST_BLOCK_ROWS = ST_THREADS * ST_ROWS_PER_THREAD;   // 1024 rows per block

__device__ __forceinline__ void decode_row(const float *__restrict__ p, uint8_t &sym, uint8_t &qual) {
    float best = p[0];
    int arg = 0;
#pragma unroll
    for (int c = 1; c < NCLS; ++c) {
        const float v = p[c];
        if (v > best) { best = v; arg = c; }                  // first maximum wins (np.argmax)
    }
    float err = fminf(fmaxf(1.0f - best, 1e-7f), 1.0f);       // labels.py:387-401 in float32
    const float l = __double2float_rn(log10((double)err));
    const float q = fminf(-10.0f * l, 70.0f);
    qual = (uint8_t)((int)q + 33);
    // '*ACGT' (labels.py:342); 0 marks a gap call so the byte doubles as the keep flag
    sym = (uint8_t)((0x5447434100ull >> (8 * arg)) & 0xff);
}

// Thread t of a block owns rows [base + 4t, base + 4t + 4): the four outputs are one 32-bit store.
__global__ void __launch_bounds__(ST_THREADS) stitch_decode_kernel(const float *__restrict__ probs, int64_t n,
                                                                   uint8_t *__restrict__ sym,
                                                                   uint8_t *__restrict__ qual,
                                                                   uint32_t *__restrict__ block_count) {
    __shared__ uint32_t warp_cnt[ST_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * ST_BLOCK_ROWS + (int64_t)threadIdx.x * ST_ROWS_PER_THREAD;
    uint32_t s4 = 0, q4 = 0, kept = 0;
    if (r0 + ST_ROWS_PER_THREAD <= n) {
        // 4 rows = 20 floats = 80 B, 16-byte aligned because r0 is a multiple of 4
        const float4 *v = reinterpret_cast<const float4 *>(probs + r0 * NCLS);
        float f[20];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float4 x = __ldcs(v + i);
            f[4 * i] = x.x; f[4 * i + 1] = x.y; f[4 * i + 2] = x.z; f[4 * i + 3] = x.w;
        }
#pragma unroll
        for (int j = 0; j < ST_ROWS_PER_THREAD; ++j) {
            uint8_t s, q;
            decode_row(f + j * NCLS, s, q);
            s4 |= (uint32_t)s << (8 * j);
            q4 |= (uint32_t)q << (8 * j);
            kept += (s != 0);
        }
        *reinterpret_cast<uint32_t *>(sym + r0) = s4;
        *reinterpret_cast<uint32_t *>(qual + r0) = q4;
    } else {
        for (int j = 0; j < ST_ROWS_PER_THREAD && r0 + j < n; ++j) {
            uint8_t s, q;
            decode_row(probs + (r0 + j) * NCLS, s, q);
            sym[r0 + j] = s;
            qual[r0 + j] = q;
            kept += (s != 0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = kept;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < ST_THREADS / 32; ++w) t += warp_cnt[w];
        block_count[blockIdx.x] = t;
    }
}

// Exclusive scan of block counts into 64-bit bases; base[n_blocks] = grand total.  One block of 1024 threads walks
// the array in 1024-entry strips carrying the running total.
__global__ void __launch_bounds__(1024) stitch_scan_kernel(const uint32_t *__restrict__ block_count, int64_t n_blocks,
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
            warp_sum[lane] = w;                                // inclusive over warps
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

__global__ void __launch_bounds__(ST_THREADS) stitch_scatter_kernel(const uint8_t *__restrict__ sym,
                                                                    const uint8_t *__restrict__ qual, int64_t n,
                                                                    const int64_t *__restrict__ block_base,
                                                                    const int64_t *__restrict__ seg_base,
                                                                    int64_t n_seg, uint8_t *__restrict__ seq_out,
                                                                    uint8_t *__restrict__ qual_out,
                                                                    int64_t *__restrict__ seg_out_off) {
    __shared__ uint32_t warp_cnt[ST_THREADS / 32];
    const int64_t r0 = (int64_t)blockIdx.x * ST_BLOCK_ROWS + (int64_t)threadIdx.x * ST_ROWS_PER_THREAD;
    uint32_t s4 = 0, q4 = 0;
    if (r0 + ST_ROWS_PER_THREAD <= n) {
        s4 = *reinterpret_cast<const uint32_t *>(sym + r0);
        q4 = *reinterpret_cast<const uint32_t *>(qual + r0);
    } else {
        for (int j = 0; j < ST_ROWS_PER_THREAD && r0 + j < n; ++j) {
            s4 |= (uint32_t)sym[r0 + j] << (8 * j);
            q4 |= (uint32_t)qual[r0 + j] << (8 * j);
        }
    }
    uint32_t mine = 0;
#pragma unroll
    for (int j = 0; j < ST_ROWS_PER_THREAD; ++j) mine += ((s4 >> (8 * j)) & 0xff) != 0;
    // exclusive rank of this thread's first survivor inside the block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_cnt[warp] = x;
    __syncthreads();
    uint32_t before = x - mine;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    int64_t o = block_base[blockIdx.x] + before;
    const int64_t o_first = o;
#pragma unroll
    for (int j = 0; j < ST_ROWS_PER_THREAD; ++j) {
        const uint8_t s = (s4 >> (8 * j)) & 0xff;
        if (s) {
            seq_out[o] = s;
            if (qual_out) qual_out[o] = (q4 >> (8 * j)) & 0xff;
            ++o;
        }
    }
    // range starts: the ranges are sorted and non-empty, so the ones beginning inside this thread's four rows are a
    // contiguous slice of seg_base found by one lower_bound
    if (r0 < n) {
        int64_t lo = 0, hi = n_seg;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (seg_base[mid] < r0) lo = mid + 1; else hi = mid;
        }
        for (int64_t k = lo; k < n_seg && seg_base[k] < r0 + ST_ROWS_PER_THREAD; ++k) {
            const int j_start = (int)(seg_base[k] - r0);
            int64_t off = o_first;
            for (int j = 0; j < j_start; ++j) off += ((s4 >> (8 * j)) & 0xff) != 0;
            seg_out_off[k] = off;
        }
    }
}

}  // namespace

// d_probs: [n][5] float32, the kept row ranges back to back; seg_base_dev: [n_seg] first row of each range (sorted,
// ranges non-empty); scratch: 2n bytes + (n_blocks) u32 + (n_blocks + 1) i64.
int stitch_dev(const float *d_probs, int64_t n, const int64_t *seg_base_dev, int64_t n_seg, uint8_t *d_seq,
               uint8_t *d_qual, int64_t *d_seg_out_off, uint8_t *scratch, cudaStream_t stream) {
    const int64_t n_blocks = (n + ST_BLOCK_ROWS - 1) / ST_BLOCK_ROWS;
    const int64_t n4 = (n + 3) & ~(int64_t)3;
    uint8_t *d_sym = scratch;
    uint8_t *d_q = scratch + n4;
    size_t off = (size_t)(2 * n4 + 15) & ~(size_t)15;
    uint32_t *d_cnt = reinterpret_cast<uint32_t *>(scratch + off);
    off = (off + (size_t)n_blocks * 4 + 15) & ~(size_t)15;
    int64_t *d_base = reinterpret_cast<int64_t *>(scratch + off);
    stitch_decode_kernel<<<(unsigned)n_blocks, ST_THREADS, 0, stream>>>(d_probs, n, d_sym, d_q, d_cnt);
    stitch_scan_kernel<<<1, 1024, 0, stream>>>(d_cnt, n_blocks, d_base);
    stitch_scatter_kernel<<<(unsigned)n_blocks, ST_THREADS, 0, stream>>>(d_sym, d_q, n, d_base, seg_base_dev, n_seg,
                                                                         d_seq, d_qual, d_seg_out_off);
    MDK_CUDA(cudaGetLastError());
    // total survivors -> seg_out_off[n_seg]
    MDK_CUDA(cudaMemcpyAsync(d_seg_out_off + n_seg, d_base + n_blocks, sizeof(int64_t), cudaMemcpyDeviceToDevice,
                             stream));
    return MDK_OK;
}

size_t stitch_scratch_bytes(int64_t n) {
    const int64_t n_blocks = (n + ST_BLOCK_ROWS - 1) / ST_BLOCK_ROWS;
    const int64_t n4 = (n + 3) & ~(int64_t)3;
    return (size_t)(2 * n4) + 16 + (size_t)n_blocks * 4 + 16 + (size_t)(n_blocks + 1) * 8 + 16;
}

}  // namespace mdk

using namespace mdk;

extern "C" {

int mdk_stitch_consensus_dev(int device, const float *probs_dev, int64_t n_rows, const int64_t *seg_base,
                             int64_t n_seg, uint8_t *seq_out_dev, uint8_t *qual_out_dev, int64_t *seg_out_off) {
    MDK_REQUIRE(n_rows >= 0 && n_seg >= 0, MDK_ERR_ARG, "stitch_consensus: negative size");
    MDK_REQUIRE(seg_out_off, MDK_ERR_ARG, "stitch_consensus: NULL seg_out_off");
    if (n_rows == 0 || n_seg == 0) {
        MDK_REQUIRE(n_rows == 0 && n_seg == 0, MDK_ERR_ARG, "stitch_consensus: rows without ranges (or vice versa)");
        seg_out_off[0] = 0;
        return MDK_OK;
    }
    MDK_REQUIRE(probs_dev && seg_base && seq_out_dev, MDK_ERR_ARG, "stitch_consensus: NULL pointer");
    MDK_REQUIRE(seg_base[0] == 0, MDK_ERR_ARG, "stitch_consensus: first range must start at row 0");
    for (int64_t k = 1; k < n_seg; ++k)
        MDK_REQUIRE(seg_base[k] > seg_base[k - 1] && seg_base[k] < n_rows, MDK_ERR_ARG,
                    "stitch_consensus: range starts must be strictly increasing and < n_rows");
    MDK_CUDA(cudaSetDevice(device));
    uint8_t *buf = nullptr;
    const size_t b_seg = ((size_t)n_seg * 8 + 15) & ~(size_t)15, b_off = ((size_t)(n_seg + 1) * 8 + 15) & ~(size_t)15;
    // per-host-thread cached scratch: a cudaMalloc / cudaFree pair per call costs more than the kernels of a contig
    MDK_CUDA(plp_scratch(b_seg + b_off + stitch_scratch_bytes(n_rows), &buf, 2));
    int64_t *d_seg = reinterpret_cast<int64_t *>(buf), *d_off = reinterpret_cast<int64_t *>(buf + b_seg);
    cudaError_t err = cudaMemcpyAsync(d_seg, seg_base, (size_t)n_seg * 8, cudaMemcpyHostToDevice, 0);
    int rc = MDK_OK;
    if (err == cudaSuccess)
        rc = stitch_dev(probs_dev, n_rows, d_seg, n_seg, seq_out_dev, qual_out_dev, d_off, buf + b_seg + b_off, 0);
    if (rc == MDK_OK && err == cudaSuccess)
        err = cudaMemcpy(seg_out_off, d_off, (size_t)(n_seg + 1) * 8, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return cuda_fail(err, "stitch_consensus_dev", __FILE__, __LINE__);
    return rc;
}

int mdk_stitch_consensus(int device, const float *const *seg_probs, const int64_t *seg_rows, int64_t n_seg,
                         uint8_t *seq_out, uint8_t *qual_out, int64_t *seg_out_off) {
    MDK_REQUIRE(n_seg >= 0, MDK_ERR_ARG, "stitch_consensus: n_seg < 0");
    MDK_REQUIRE(seg_out_off, MDK_ERR_ARG, "stitch_consensus: NULL seg_out_off");
    if (n_seg == 0) { seg_out_off[0] = 0; return MDK_OK; }
    MDK_REQUIRE(seg_probs && seg_rows && seq_out, MDK_ERR_ARG, "stitch_consensus: NULL pointer");
    int64_t n = 0;
    std::vector<int64_t> base((size_t)n_seg);
    for (int64_t k = 0; k < n_seg; ++k) {
        MDK_REQUIRE(seg_rows[k] > 0 && seg_probs[k], MDK_ERR_ARG,
                    "stitch_consensus: every range needs rows > 0 and a probabilities pointer");
        base[(size_t)k] = n;
        n += seg_rows[k];
    }
    MDK_CUDA(cudaSetDevice(device));
    uint8_t *buf = nullptr;
    const size_t b_probs = ((size_t)n * NCLS * 4 + 15) & ~(size_t)15;
    const size_t b_out = ((size_t)n + 15) & ~(size_t)15;
    MDK_CUDA(plp_scratch(b_probs + 2 * b_out, &buf, 3));
    float *d_probs = reinterpret_cast<float *>(buf);
    uint8_t *d_seq = buf + b_probs, *d_qual = d_seq + b_out;
    cudaError_t err = cudaSuccess;
    for (int64_t k = 0; k < n_seg && err == cudaSuccess; ++k)
        err = cudaMemcpyAsync(d_probs + base[(size_t)k] * NCLS, seg_probs[k], (size_t)seg_rows[k] * NCLS * 4,
                              cudaMemcpyHostToDevice, 0);
    int rc = MDK_OK;
    if (err == cudaSuccess)
        rc = mdk_stitch_consensus_dev(device, d_probs, n, base.data(), n_seg, d_seq, qual_out ? d_qual : nullptr,
                                      seg_out_off);
    if (rc == MDK_OK && err == cudaSuccess) {
        const size_t total = (size_t)seg_out_off[n_seg];
        if (total) {
            err = cudaMemcpy(seq_out, d_seq, total, cudaMemcpyDeviceToHost);
            if (err == cudaSuccess && qual_out) err = cudaMemcpy(qual_out, d_qual, total, cudaMemcpyDeviceToHost);
        }
    }
    if (err != cudaSuccess) return cuda_fail(err, "stitch_consensus", __FILE__, __LINE__);
    return rc;
}

}  // extern "C"
