// Pileup-counts featuriser on the GPU: the per-base work of calculate_pileup (src/medaka_counts.c:199-372)
// over BAM-packed alignment records (32-bit CIGAR ops, 4-bit sequence), without htslib.
//
// htslib's pileup walks reference positions and, per position, the reads covering it.  Here the loop nest is
// inverted so that it is data-parallel over reads and CIGAR operations:
//   1. plp_walk_kernel   (WARP per read, lanes over CIGAR ops, 32 at a time): reference / query cursor of every op by
//                        warp prefix sums; how insertion runs attach to the reference base before them (htslib's
//                        "peek the next operation" rule, restated in oracle/pileup_oracle.py and pinned on the
//                        reference's real-BAM regression numbers) by a segmented warp scan; the read's coverage as
//                        two entries of a difference array (+1 at its first position in the region, -1 behind its
//                        last); the longest insertion behind every position (atomicMax, insertions only).
//   2. block scans       depth = prefix sum of the difference array; width[pos] = (depth > 0) + longest insertion;
//                        first column of every position = exclusive prefix sum of width.  Three-phase scans (per-block
//                        sums, one small block over the block sums, per-block apply) - no single-block pass over the
//                        region; the (major, minor) arrays come out of the last phase (medaka_counts.c:274-277).
//   3. plp_count_kernel  (thread per op, the warp gangs up on long ops): counts[col][dtype*10 + base] += 1 in the
//                        'acgtACGTdD' feature order (medaka_counts.h:19-30), deletions at minor 0, inserted bases at
//                        minors 1..k (medaka_counts.c:314-357).  32-bit reductions (RED.ADD.U32); widened to the
//                        reference's size_t matrix at the end.
// Read filter: flags and mapQ on the device (medaka_bamiter.c:19-21); tag / read-group / datatype resolution is
// done by the host reader (medaka_b200/bam.py), which hands over a per-read dtype index.
// Reference-skip (N) operations: the read covers the skipped positions (a column exists, nothing is counted there,
// medaka_counts.c:282), and an insertion right behind a skip widens its position's column group but its bases are
// NOT counted (the `continue` at :282 comes before the base loop; the max_ins loop at :259-263 does not look at it).
// All of it is HBM-bound integer/byte work: no tensor cores.  The device scratch is cached per host thread (no
// cudaMalloc on the steady-state path) and nothing synchronises with the host before the final copies.
#include "common.cuh"

namespace mdk {

constexpr int PLP_FILTER_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800;   // UNMAP|SECONDARY|QCFAIL|DUP|SUPPLEMENTARY
// 4-bit IUPAC code (+16 if reverse strand) -> index in 'acgtACGTdD' (src/medaka_counts.h:25-30)
__constant__ int8_t c_num2countbase[32] = {-1, 4, 5, -1, 6, -1, -1, -1, 7, -1, -1, -1, -1, -1, -1, -1,
                                           -1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
constexpr int OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_P = 6, OP_EQ = 7, OP_X = 8;
constexpr uint32_t INS_NOCOUNT = 0x80000000u;   // op_ins flag: the run hangs off a reference skip

__device__ __forceinline__ bool read_passes(uint16_t flag, uint8_t mapq, int min_mapq) {
    return !(flag & PLP_FILTER_FLAGS) && (int)mapq >= min_mapq;
}
__device__ __forceinline__ bool consumes_ref(int op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool is_match(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

// State of the insertion-run scan after an op: kind of the last op that is neither I nor P (0 none yet, 1 consumes the
// reference, 2 reference skip, 3 anything else) and the inserted bases seen since then.  Associative combine: a later
// segment that contains such an op overrides the kind and restarts the sum.
struct RunState {
    uint32_t kind;
    uint32_t ins;
};
__device__ __forceinline__ RunState run_combine(RunState a, RunState b) {
    RunState r;
    r.kind = b.kind ? b.kind : a.kind;
    r.ins = b.kind ? b.ins : a.ins + b.ins;
    return r;
}

// op_ref[k] : reference cursor at the start of op k;  op_qry[k] : query cursor at the start of op k
// op_ins[k] : for an I op attached to reference position op_ref[k] - 1: 1 + the inserted bases before it in its run
//             (| INS_NOCOUNT when the run follows a reference skip); 0 otherwise
// Coverage and longest insertions are only recorded for reads that pass the filter.
__global__ void __launch_bounds__(256) plp_walk_kernel(int64_t n_rec, const int32_t *__restrict__ pos,
                                                       const uint16_t *__restrict__ flag, const uint8_t *__restrict__ mapq,
                                                       int min_mapq, const uint32_t *__restrict__ cigar,
                                                       const int64_t *__restrict__ cigar_off, int32_t start, int32_t end,
                                                       int32_t *__restrict__ op_ref, int32_t *__restrict__ op_qry,
                                                       uint32_t *__restrict__ op_ins, int32_t *__restrict__ cov,
                                                       int32_t *__restrict__ maxins) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rec) return;
    const bool pass = read_passes(flag[r], mapq[r], min_mapq);
    const int64_t k0 = cigar_off[r], k1 = cigar_off[r + 1];
    int32_t x = pos[r], y = 0;                 // cursors at the start of the current chunk of 32 ops
    RunState carry{0u, 0u};
    for (int64_t kb = k0; kb < k1; kb += 32) {
        const int64_t k = kb + lane;
        const bool live = k < k1;
        const uint32_t c = live ? cigar[k] : 0u;
        const int op = live ? (int)(c & 0xF) : OP_P;     // padding consumes nothing and breaks nothing
        const int32_t len = live ? (int32_t)(c >> 4) : 0;
        int32_t dx = consumes_ref(op) ? len : 0;
        int32_t dy = (is_match(op) || op == OP_I || op == OP_S) ? len : 0;
        RunState st;
        st.kind = (op == OP_I || op == OP_P) ? 0u : (op == OP_N ? 2u : (consumes_ref(op) ? 1u : 3u));
        st.ins = op == OP_I ? (uint32_t)len : 0u;
        // inclusive warp scans
        int32_t sx = dx, sy = dy;
        RunState ss = st;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t tx = __shfl_up_sync(0xffffffffu, sx, o), ty = __shfl_up_sync(0xffffffffu, sy, o);
            RunState t;
            t.kind = __shfl_up_sync(0xffffffffu, ss.kind, o);
            t.ins = __shfl_up_sync(0xffffffffu, ss.ins, o);
            if (lane >= o) {
                sx += tx;
                sy += ty;
                ss = run_combine(t, ss);
            }
        }
        // state BEFORE this op = carry (+) inclusive state of the previous lane
        RunState before;
        before.kind = __shfl_up_sync(0xffffffffu, ss.kind, 1);
        before.ins = __shfl_up_sync(0xffffffffu, ss.ins, 1);
        if (lane == 0) before = RunState{0u, 0u};
        before = run_combine(carry, before);
        if (live) {
            const int32_t xr = x + sx - dx, yq = y + sy - dy;
            op_ref[k] = xr;
            op_qry[k] = yq;
            uint32_t ins = 0;
            if (op == OP_I && (before.kind == 1u || before.kind == 2u)) {
                ins = (before.ins + 1u) | (before.kind == 2u ? INS_NOCOUNT : 0u);
                const int32_t p = xr - 1;
                if (pass && p >= start && p < end) atomicMax(&maxins[p - start], (int32_t)(before.ins + (uint32_t)len));
            }
            op_ins[k] = ins;
        }
        x += __shfl_sync(0xffffffffu, sx, 31);
        y += __shfl_sync(0xffffffffu, sy, 31);
        RunState last;
        last.kind = __shfl_sync(0xffffffffu, ss.kind, 31);
        last.ins = __shfl_sync(0xffffffffu, ss.ins, 31);
        carry = run_combine(carry, last);
    }
    // the read covers [pos, x): htslib's pileup lists it at every one of those positions (deleted and skipped ones too)
    if (lane == 0 && pass) {
        const int32_t lo = max(pos[r], start), hi = min(x, end);
        if (lo < hi) {
            atomicAdd(&cov[lo - start], 1);
            atomicAdd(&cov[hi - start], -1);
        }
    }
}

// ---------------------------------------------------------------------------------------------- three-phase scans
constexpr int SC_THREADS = 256, SC_PER_THREAD = 8, SC_BLOCK = SC_THREADS * SC_PER_THREAD;   // 2048 positions per block

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *total) {
    __shared__ int64_t wsum[SC_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int64_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < SC_THREADS / 32; ++w) {
        if (w < warp) before += wsum[w];
        all += wsum[w];
    }
    if (total) *total = all;
    return before + x - v;
}

// phase 1: per-block sum of the coverage difference array
__global__ void __launch_bounds__(SC_THREADS) plp_sum_cov_kernel(int32_t L, const int32_t *__restrict__ cov,
                                                                 int64_t *__restrict__ blk) {
    const int32_t i0 = blockIdx.x * SC_BLOCK + threadIdx.x * SC_PER_THREAD;
    int64_t s = 0;
#pragma unroll
    for (int j = 0; j < SC_PER_THREAD; ++j)
        if (i0 + j < L) s += cov[i0 + j];
    int64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// phase 2: exclusive scan of up to a few thousand block sums by one block; blk[n] = grand total
__global__ void __launch_bounds__(SC_THREADS) plp_scan_blocks_kernel(int64_t n, int64_t *__restrict__ blk) {
    __shared__ int64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t s = 0; s < n; s += SC_THREADS) {
        const int64_t i = s + threadIdx.x;
        const int64_t v = i < n ? blk[i] : 0;
        int64_t total;
        const int64_t ex = block_exclusive_scan(v, &total);
        const int64_t c = carry;
        if (i < n) blk[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) blk[n] = carry;
}

// phase 3: depth -> width (kept in `maxins`' array) -> per-block sum of width
__global__ void __launch_bounds__(SC_THREADS) plp_width_kernel(int32_t L, const int32_t *__restrict__ cov,
                                                               const int64_t *__restrict__ blk_cov,
                                                               int32_t *__restrict__ width, int64_t *__restrict__ blk_w) {
    const int32_t i0 = blockIdx.x * SC_BLOCK + threadIdx.x * SC_PER_THREAD;
    int32_t d[SC_PER_THREAD];
    int64_t s = 0;
#pragma unroll
    for (int j = 0; j < SC_PER_THREAD; ++j) {
        d[j] = i0 + j < L ? cov[i0 + j] : 0;
        s += d[j];
    }
    int64_t depth = blk_cov[blockIdx.x] + block_exclusive_scan(s, nullptr);
    int64_t w = 0;
#pragma unroll
    for (int j = 0; j < SC_PER_THREAD; ++j) {
        depth += d[j];
        if (i0 + j < L) {
            const int32_t wj = depth > 0 ? 1 + width[i0 + j] : 0;     // (width[] holds the longest insertion so far)
            width[i0 + j] = wj;
            w += wj;
        }
    }
    int64_t total;
    block_exclusive_scan(w, &total);
    if (threadIdx.x == 0) blk_w[blockIdx.x] = total;
}

// phase 5 (after the block sums of width have been scanned): first column of every position, position arrays
__global__ void __launch_bounds__(SC_THREADS) plp_columns_kernel(int32_t L, int32_t start, const int32_t *__restrict__ width,
                                                                 const int64_t *__restrict__ blk_w, int64_t max_cols,
                                                                 int64_t *__restrict__ col_off, int64_t *__restrict__ major,
                                                                 int64_t *__restrict__ minor) {
    const int32_t i0 = blockIdx.x * SC_BLOCK + threadIdx.x * SC_PER_THREAD;
    int32_t w[SC_PER_THREAD];
    int64_t s = 0;
#pragma unroll
    for (int j = 0; j < SC_PER_THREAD; ++j) {
        w[j] = i0 + j < L ? width[i0 + j] : 0;
        s += w[j];
    }
    int64_t c = blk_w[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
    for (int j = 0; j < SC_PER_THREAD; ++j) {
        if (i0 + j < L) {
            col_off[i0 + j] = c;
            for (int32_t m = 0; m < w[j]; ++m) {
                if (c + m < max_cols) {
                    major[c + m] = (int64_t)start + i0 + j;
                    minor[c + m] = m;
                }
            }
            c += w[j];
        }
    }
}

// ---------------------------------------------------------------------------------------------- counting
__device__ __forceinline__ int seq_code(const uint8_t *__restrict__ seq, int64_t base, int32_t q) {
    const uint8_t b = seq[base + (q >> 1)];
    return (q & 1) ? (b & 0xF) : (b >> 4);
}

struct CountOp {
    int32_t lo, hi;        // reference positions [lo, hi) (matches, deletions) or inserted-base indices [0, hi) (lo = 0)
    int32_t x0, q0;        // reference / query cursor at the start of the op
    int64_t col0;          // insertions: first column of the op's bases
    int64_t sbase;         // byte offset of the read's packed sequence
    int32_t fbase;         // 10 * dtype
    int32_t rev;           // 16 when the read is on the reverse strand
    int32_t kind;          // 0 nothing, 1 match, 2 deletion, 3 insertion
};

__device__ __forceinline__ void count_one(const CountOp &o, int32_t i, const uint8_t *__restrict__ seq,
                                          const int64_t *__restrict__ col_off, int32_t start, int F, int64_t max_cols,
                                          uint32_t *__restrict__ counts) {
    if (o.kind == 1) {
        const int bi = c_num2countbase[seq_code(seq, o.sbase, o.q0 + (i - o.x0)) + o.rev];
        const int64_t col = col_off[i - start];
        if (bi >= 0 && col < max_cols) atomicAdd(&counts[col * F + o.fbase + bi], 1u);
    } else if (o.kind == 2) {
        const int64_t col = col_off[i - start];
        if (col < max_cols) atomicAdd(&counts[col * F + o.fbase + (o.rev ? 8 : 9)], 1u);   // rev_del / fwd_del
    } else {
        const int bi = c_num2countbase[seq_code(seq, o.sbase, o.q0 + i) + o.rev];
        if (bi >= 0 && o.col0 + i < max_cols) atomicAdd(&counts[(o.col0 + i) * F + o.fbase + bi], 1u);
    }
}

constexpr int CNT_INLINE = 12;     // bases an op's own thread handles; longer ops are shared out over the warp

__global__ void __launch_bounds__(256) plp_count_kernel(int64_t n_ops, const int32_t *__restrict__ op_rec,
                                                        const uint32_t *__restrict__ cigar, const int32_t *__restrict__ op_ref,
                                                        const int32_t *__restrict__ op_qry, const uint32_t *__restrict__ op_ins,
                                                        const uint16_t *__restrict__ flag, const uint8_t *__restrict__ mapq,
                                                        const uint8_t *__restrict__ dtype, const uint8_t *__restrict__ seq,
                                                        const int64_t *__restrict__ seq_off, int min_mapq, int32_t start,
                                                        int32_t end, int num_dtypes, const int64_t *__restrict__ col_off,
                                                        int64_t max_cols, uint32_t *__restrict__ counts) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int F = 10 * num_dtypes;
    CountOp o;
    o.kind = 0;
    o.lo = o.hi = 0;
    if (k < n_ops) {
        const int r = op_rec[k];
        const uint16_t fl = flag[r];
        if (read_passes(fl, mapq[r], min_mapq)) {
            const uint32_t c = cigar[k];
            const int op = c & 0xF;
            const int32_t len = (int32_t)(c >> 4);
            o.x0 = op_ref[k];
            o.q0 = op_qry[k];
            o.sbase = seq_off[r];
            o.fbase = 10 * (int)dtype[r];
            o.rev = (fl & 0x10) ? 16 : 0;
            if (is_match(op) || op == OP_D) {
                o.kind = is_match(op) ? 1 : 2;
                o.lo = max(o.x0, start);
                o.hi = min(o.x0 + len, end);
            } else if (op == OP_I) {
                const uint32_t ins = op_ins[k];
                const int32_t p = o.x0 - 1;
                if (ins != 0u && !(ins & INS_NOCOUNT) && p >= start && p < end) {
                    o.kind = 3;
                    o.lo = 0;
                    o.hi = len;
                    o.col0 = col_off[p - start] + (int64_t)(ins & ~INS_NOCOUNT);   // minor 1 + bases before it
                }
            }
            if (o.hi <= o.lo) o.kind = 0;
        }
    }
    const bool big = o.kind != 0 && o.hi - o.lo > CNT_INLINE;
    if (o.kind != 0 && !big)
        for (int32_t i = o.lo; i < o.hi; ++i) count_one(o, i, seq, col_off, start, F, max_cols, counts);
    // long operations (a long match, a multi-kilobase deletion): one after the other, 32 bases per step
    uint32_t todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        CountOp w;
        w.lo = __shfl_sync(0xffffffffu, o.lo, src);
        w.hi = __shfl_sync(0xffffffffu, o.hi, src);
        w.x0 = __shfl_sync(0xffffffffu, o.x0, src);
        w.q0 = __shfl_sync(0xffffffffu, o.q0, src);
        w.col0 = __shfl_sync(0xffffffffu, o.col0, src);
        w.sbase = __shfl_sync(0xffffffffu, o.sbase, src);
        w.fbase = __shfl_sync(0xffffffffu, o.fbase, src);
        w.rev = __shfl_sync(0xffffffffu, o.rev, src);
        w.kind = __shfl_sync(0xffffffffu, o.kind, src);
        for (int32_t i = w.lo + lane; i < w.hi; i += 32) count_one(w, i, seq, col_off, start, F, max_cols, counts);
    }
}

__global__ void plp_op_rec_kernel(int64_t n_rec, const int64_t *__restrict__ cigar_off, int32_t *__restrict__ op_rec) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rec) return;
    for (int64_t k = cigar_off[r] + lane; k < cigar_off[r + 1]; k += 32) op_rec[k] = (int32_t)r;
}

__global__ void plp_widen_kernel(int64_t n, const uint32_t *__restrict__ src, uint64_t *__restrict__ dst) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// per-host-thread device scratch (grown on demand, reused across calls)
struct PlpScratch {
    int device = -1;
    uint8_t *buf = nullptr;
    size_t cap = 0;
    ~PlpScratch() {
        if (buf) cudaFree(buf);
    }
};
static thread_local PlpScratch g_plp_scratch[4];      // 0: kernel scratch, 1: staging of the host-buffer entry point,
                                                       // 2 / 3: the stitch entry points (stitch.cu)

cudaError_t plp_scratch(size_t bytes, uint8_t **out, int slot) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    PlpScratch &s = g_plp_scratch[slot];
    if (s.device != dev || s.cap < bytes) {
        if (s.buf) cudaFree(s.buf);
        s.buf = nullptr;
        s.cap = 0;
        const size_t want = bytes + bytes / 4 + (1 << 20);
        e = cudaMalloc(&s.buf, want);
        if (e != cudaSuccess) return e;
        s.cap = want;
        s.device = dev;
    }
    *out = s.buf;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------
// Host driver (device pointers in, device pointers out).  Returns the number of columns through *n_cols_host;
// if it exceeds max_cols the outputs are incomplete and the caller re-runs with a larger buffer (the reference's
// enlarge_plp_data, medaka_counts.c:266-271).
int pileup_counts_dev(int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                      const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, int64_t n_ops,
                      const uint8_t *seq, const int64_t *seq_off, int32_t start, int32_t end, int num_dtypes,
                      int min_mapq, int64_t max_cols, uint64_t *counts, int64_t *major, int64_t *minor,
                      int64_t *n_cols_host, cudaStream_t s) {
    const int32_t L = end - start;
    *n_cols_host = 0;
    if (L <= 0 || n_rec == 0 || n_ops == 0) return MDK_OK;
    const int F = 10 * num_dtypes;
    const int64_t n_blk = (L + SC_BLOCK - 1) / SC_BLOCK;
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_rec = take((size_t)n_ops * 4), o_ref = take((size_t)n_ops * 4), o_qry = take((size_t)n_ops * 4),
                 o_ins = take((size_t)n_ops * 4), o_cov = take((size_t)(L + 1) * 4), o_w = take((size_t)L * 4),
                 o_col = take((size_t)L * 8), o_bc = take((size_t)(n_blk + 1) * 8), o_bw = take((size_t)(n_blk + 1) * 8),
                 o_cnt = take((size_t)max_cols * F * 4);
    uint8_t *scratch = nullptr;
    MDK_CUDA(plp_scratch(off, &scratch, 0));
    int32_t *op_rec = (int32_t *)(scratch + o_rec), *op_ref = (int32_t *)(scratch + o_ref), *op_qry = (int32_t *)(scratch + o_qry);
    uint32_t *op_ins = (uint32_t *)(scratch + o_ins);
    int32_t *cov = (int32_t *)(scratch + o_cov), *width = (int32_t *)(scratch + o_w);
    int64_t *col_off = (int64_t *)(scratch + o_col), *blk_cov = (int64_t *)(scratch + o_bc), *blk_w = (int64_t *)(scratch + o_bw);
    uint32_t *cnt32 = (uint32_t *)(scratch + o_cnt);
    // cov and width are adjacent: one memset
    MDK_CUDA(cudaMemsetAsync(scratch + o_cov, 0, (o_w - o_cov) + (size_t)L * 4, s));
    if (max_cols > 0) MDK_CUDA(cudaMemsetAsync(cnt32, 0, (size_t)max_cols * F * 4, s));
    const unsigned wb = (unsigned)((n_rec * 32 + 255) / 256), ob = (unsigned)((n_ops + 255) / 256);
    plp_op_rec_kernel<<<wb, 256, 0, s>>>(n_rec, cigar_off, op_rec);
    plp_walk_kernel<<<wb, 256, 0, s>>>(n_rec, pos, flag, mapq, min_mapq, cigar, cigar_off, start, end, op_ref, op_qry, op_ins,
                                       cov, width);
    plp_sum_cov_kernel<<<(unsigned)n_blk, SC_THREADS, 0, s>>>(L, cov, blk_cov);
    plp_scan_blocks_kernel<<<1, SC_THREADS, 0, s>>>(n_blk, blk_cov);
    plp_width_kernel<<<(unsigned)n_blk, SC_THREADS, 0, s>>>(L, cov, blk_cov, width, blk_w);
    plp_scan_blocks_kernel<<<1, SC_THREADS, 0, s>>>(n_blk, blk_w);
    plp_columns_kernel<<<(unsigned)n_blk, SC_THREADS, 0, s>>>(L, start, width, blk_w, max_cols, col_off, major, minor);
    if (max_cols > 0) {
        plp_count_kernel<<<ob, 256, 0, s>>>(n_ops, op_rec, cigar, op_ref, op_qry, op_ins, flag, mapq, dtype, seq, seq_off, min_mapq,
                                            start, end, num_dtypes, col_off, max_cols, cnt32);
        const int64_t n_out = max_cols * F;
        plp_widen_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, s>>>(n_out, cnt32, counts);
    }
    MDK_CUDA(cudaGetLastError());
    MDK_CUDA(cudaMemcpyAsync(n_cols_host, blk_w + n_blk, 8, cudaMemcpyDeviceToHost, s));
    MDK_CUDA(cudaStreamSynchronize(s));
    return MDK_OK;
}

}  // namespace mdk
