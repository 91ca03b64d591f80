// Pileup-counts featuriser on the GPU: the per-base work of calculate_pileup (src/medaka_counts.c:199-372)
// over BAM-packed alignment records (32-bit CIGAR ops, 4-bit sequence), without htslib.
//
// htslib's pileup walks reference positions and, per position, the reads covering it.  Here the loop nest is
// inverted so that it is data-parallel over CIGAR operations:
//   1. plp_walk_kernel   (thread per read): running reference / query cursors of every CIGAR op, and how
//                        insertion runs attach to the reference base before them (htslib's "peek the next
//                        operation" rule, restated in oracle/pileup_oracle.py and pinned on the reference's real
//                        BAM regression numbers);
//   2. plp_cover_kernel  (thread per op): per reference position, width[pos] = 1 + longest insertion after it,
//                        0 where no passing read covers it (atomicMax);
//   3. plp_scan_kernel   : exclusive scan of width -> first column of every position, total column count;
//                        writes the (major, minor) position arrays (medaka_counts.c:274-277);
//   4. plp_count_kernel  (thread per op): counts[col][dtype*10 + base] += 1 with the 'acgtACGTdD' feature order
//                        (medaka_counts.h:19-30), deletions at minor 0, inserted bases at minors 1..k
//                        (medaka_counts.c:314-357); 64-bit atomics.
// Read filter: flags and mapQ on the device (medaka_bamiter.c:19-21); tag / read-group / datatype resolution is
// done by the host reader (medaka_b200/bam.py), which hands over a per-read dtype index.
// All of it is HBM-bound integer/byte work: no tensor cores.
#include "common.cuh"

namespace mdk {

constexpr int PLP_FILTER_FLAGS = 0x4 | 0x100 | 0x200 | 0x400 | 0x800;   // UNMAP|SECONDARY|QCFAIL|DUP|SUPPLEMENTARY
// 4-bit IUPAC code (+16 if reverse strand) -> index in 'acgtACGTdD' (src/medaka_counts.h:25-30)
__constant__ int8_t c_num2countbase[32] = {-1, 4, 5, -1, 6, -1, -1, -1, 7, -1, -1, -1, -1, -1, -1, -1,
                                           -1, 0, 1, -1, 2, -1, -1, -1, 3, -1, -1, -1, -1, -1, -1, -1};
constexpr int OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3, OP_S = 4, OP_P = 6, OP_EQ = 7, OP_X = 8;

__device__ __forceinline__ bool read_passes(uint16_t flag, uint8_t mapq, int min_mapq) {
    return !(flag & PLP_FILTER_FLAGS) && (int)mapq >= min_mapq;
}
__device__ __forceinline__ bool consumes_ref(int op) { return op == OP_M || op == OP_D || op == OP_N || op == OP_EQ || op == OP_X; }
__device__ __forceinline__ bool is_match(int op) { return op == OP_M || op == OP_EQ || op == OP_X; }

// op_ref[k]   : reference cursor at the start of op k
// op_qry[k]   : query cursor at the start of op k
// op_minor[k] : for I ops, the minor index of the op's first base minus 1 (0 for the first I of a run) when the run
//               is attached to reference position op_ref[k]-1, or -1 when it has no reference base before it
// ins_total[k]: for the FIRST I op of an attached run, the total inserted length of the run (else 0)
__global__ void plp_walk_kernel(int64_t n_rec, const int32_t *__restrict__ pos, const uint32_t *__restrict__ cigar,
                                const int64_t *__restrict__ cigar_off, int32_t *__restrict__ op_ref,
                                int32_t *__restrict__ op_qry, int32_t *__restrict__ op_minor,
                                int32_t *__restrict__ ins_total) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    int32_t x = pos[r], y = 0;
    bool have_ref = false;     // a reference-consuming op precedes (directly, or through I/P ops)
    bool adjacent = false;     // ... with nothing but I / P ops in between
    int32_t run = 0;           // inserted bases so far in the current attached run
    int64_t run_first = -1;
    for (int64_t k = cigar_off[r]; k < cigar_off[r + 1]; ++k) {
        const uint32_t c = cigar[k];
        const int op = c & 0xF;
        const int32_t len = (int32_t)(c >> 4);
        op_ref[k] = x;
        op_qry[k] = y;
        op_minor[k] = -1;
        ins_total[k] = 0;
        if (op == OP_I) {
            if (have_ref && adjacent) {
                op_minor[k] = run;
                if (run_first < 0) run_first = k;
                run += len;
                ins_total[run_first] = run;
            }
            y += len;
        } else if (op == OP_P) {
            // padding neither consumes anything nor breaks an insertion run (htslib skips it when peeking)
        } else {
            adjacent = false;
            run = 0;
            run_first = -1;
            if (consumes_ref(op)) {
                x += len;
                have_ref = true;
                adjacent = true;
            }
            if (is_match(op) || op == OP_S) y += len;
        }
    }
}

__global__ void plp_cover_kernel(int64_t n_ops, const int32_t *__restrict__ op_rec, const uint32_t *__restrict__ cigar,
                                 const int32_t *__restrict__ op_ref, const int32_t *__restrict__ ins_total,
                                 const uint16_t *__restrict__ flag, const uint8_t *__restrict__ mapq, int min_mapq,
                                 int32_t start, int32_t end, int32_t *__restrict__ width) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const int r = op_rec[k];
    if (!read_passes(flag[r], mapq[r], min_mapq)) return;
    const uint32_t c = cigar[k];
    const int op = c & 0xF;
    const int32_t len = (int32_t)(c >> 4);
    if (consumes_ref(op)) {
        const int32_t lo = max(op_ref[k], start), hi = min(op_ref[k] + len, end);
        for (int32_t p = lo; p < hi; ++p)
            if (width[p - start] < 1) atomicMax(&width[p - start], 1);
    } else if (op == OP_I && ins_total[k] > 0) {
        const int32_t p = op_ref[k] - 1;
        if (p >= start && p < end) atomicMax(&width[p - start], 1 + ins_total[k]);
    }
}

// Single-block exclusive scan (regions are at most ~1e6 positions); also emits the position arrays.
__global__ void __launch_bounds__(1024) plp_scan_kernel(int32_t L, int32_t start, const int32_t *__restrict__ width,
                                                        int64_t *__restrict__ col_off, int64_t *__restrict__ n_cols) {
    __shared__ int64_t part[1024];
    const int tid = threadIdx.x;
    const int32_t per = (L + 1023) / 1024;
    const int32_t lo = min(tid * per, L), hi = min(lo + per, L);
    int64_t s = 0;
    for (int32_t i = lo; i < hi; ++i) s += width[i];
    part[tid] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        int64_t v = (tid >= off) ? part[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int64_t run = part[tid] - s;
    for (int32_t i = lo; i < hi; ++i) {
        col_off[i] = run;
        run += width[i];
    }
    if (tid == 1023) *n_cols = part[1023];
}

__global__ void plp_positions_kernel(int32_t L, int32_t start, const int32_t *__restrict__ width,
                                     const int64_t *__restrict__ col_off, int64_t max_cols,
                                     int64_t *__restrict__ major, int64_t *__restrict__ minor) {
    const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    const int64_t c0 = col_off[i];
    for (int32_t m = 0; m < width[i]; ++m) {
        if (c0 + m < max_cols) {
            major[c0 + m] = (int64_t)start + i;
            minor[c0 + m] = m;
        }
    }
}

__device__ __forceinline__ int seq_code(const uint8_t *__restrict__ seq, int64_t base, int32_t q) {
    const uint8_t b = seq[base + (q >> 1)];
    return (q & 1) ? (b & 0xF) : (b >> 4);
}

__global__ void plp_count_kernel(int64_t n_ops, const int32_t *__restrict__ op_rec, const uint32_t *__restrict__ cigar,
                                 const int32_t *__restrict__ op_ref, const int32_t *__restrict__ op_qry,
                                 const int32_t *__restrict__ op_minor, const uint16_t *__restrict__ flag,
                                 const uint8_t *__restrict__ mapq, const uint8_t *__restrict__ dtype,
                                 const uint8_t *__restrict__ seq, const int64_t *__restrict__ seq_off, int min_mapq,
                                 int32_t start, int32_t end, int num_dtypes, const int64_t *__restrict__ col_off,
                                 unsigned long long *__restrict__ counts) {
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_ops) return;
    const int r = op_rec[k];
    const uint16_t fl = flag[r];
    if (!read_passes(fl, mapq[r], min_mapq)) return;
    const uint32_t c = cigar[k];
    const int op = c & 0xF;
    const int32_t len = (int32_t)(c >> 4);
    const int F = 10 * num_dtypes;
    const int fbase = 10 * (int)dtype[r];
    const int rev = (fl & 0x10) ? 16 : 0;
    const int64_t sbase = seq_off[r];
    if (is_match(op)) {
        const int32_t x0 = op_ref[k], q0 = op_qry[k];
        const int32_t lo = max(x0, start), hi = min(x0 + len, end);
        for (int32_t p = lo; p < hi; ++p) {
            const int bi = c_num2countbase[seq_code(seq, sbase, q0 + (p - x0)) + rev];
            if (bi >= 0) atomicAdd(&counts[col_off[p - start] * F + fbase + bi], 1ULL);
        }
    } else if (op == OP_D) {
        const int32_t x0 = op_ref[k];
        const int32_t lo = max(x0, start), hi = min(x0 + len, end);
        const int bi = rev ? 8 : 9;     // rev_del / fwd_del (medaka_counts.h:21-22)
        for (int32_t p = lo; p < hi; ++p) atomicAdd(&counts[col_off[p - start] * F + fbase + bi], 1ULL);
    } else if (op == OP_I && op_minor[k] >= 0) {
        const int32_t p = op_ref[k] - 1;
        if (p >= start && p < end) {
            const int64_t c0 = col_off[p - start] + 1 + op_minor[k];
            const int32_t q0 = op_qry[k];
            for (int32_t m = 0; m < len; ++m) {
                const int bi = c_num2countbase[seq_code(seq, sbase, q0 + m) + rev];
                if (bi >= 0) atomicAdd(&counts[(c0 + m) * F + fbase + bi], 1ULL);
            }
        }
    }
}

__global__ void plp_op_rec_kernel(int64_t n_rec, const int64_t *__restrict__ cigar_off, int32_t *__restrict__ op_rec) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rec) return;
    for (int64_t k = cigar_off[r]; k < cigar_off[r + 1]; ++k) op_rec[k] = (int32_t)r;
}

// ---------------------------------------------------------------------------------------------------------
// Host driver (device pointers in, device pointers out).  Returns the number of columns through *n_cols_host;
// if it exceeds max_cols nothing is counted and the caller re-runs with a larger buffer (the reference's
// enlarge_plp_data, medaka_counts.c:266-271).
int pileup_counts_dev(int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                      const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, int64_t n_ops,
                      const uint8_t *seq, const int64_t *seq_off, int32_t start, int32_t end, int num_dtypes,
                      int min_mapq, int64_t max_cols, uint64_t *counts, int64_t *major, int64_t *minor,
                      int64_t *n_cols_host, cudaStream_t s) {
    const int32_t L = end - start;
    *n_cols_host = 0;
    if (L <= 0 || n_rec == 0 || n_ops == 0) return MDK_OK;
    int32_t *op_rec = nullptr, *op_ref = nullptr, *op_qry = nullptr, *op_minor = nullptr, *ins_total = nullptr,
            *width = nullptr;
    int64_t *col_off = nullptr, *d_ncols = nullptr;
    uint8_t *scratch = nullptr;
    const size_t b_ops = (size_t)n_ops * 4, b_w = (size_t)L * 4, b_c = (size_t)L * 8;
    MDK_CUDA(cudaMalloc(&scratch, 5 * b_ops + b_w + b_c + 64));
    op_rec = reinterpret_cast<int32_t *>(scratch);
    op_ref = op_rec + n_ops;
    op_qry = op_ref + n_ops;
    op_minor = op_qry + n_ops;
    ins_total = op_minor + n_ops;
    width = ins_total + n_ops;
    col_off = reinterpret_cast<int64_t *>(scratch + (((5 * b_ops + b_w) + 7) / 8) * 8);
    d_ncols = col_off + L;
    cudaError_t err = cudaMemsetAsync(width, 0, b_w, s);
    const unsigned rb = (unsigned)((n_rec + 127) / 128), ob = (unsigned)((n_ops + 255) / 256);
    if (err == cudaSuccess) { plp_op_rec_kernel<<<rb, 128, 0, s>>>(n_rec, cigar_off, op_rec); err = cudaGetLastError(); }
    if (err == cudaSuccess) { plp_walk_kernel<<<rb, 128, 0, s>>>(n_rec, pos, cigar, cigar_off, op_ref, op_qry, op_minor, ins_total); err = cudaGetLastError(); }
    if (err == cudaSuccess) { plp_cover_kernel<<<ob, 256, 0, s>>>(n_ops, op_rec, cigar, op_ref, ins_total, flag, mapq, min_mapq, start, end, width); err = cudaGetLastError(); }
    if (err == cudaSuccess) { plp_scan_kernel<<<1, 1024, 0, s>>>(L, start, width, col_off, d_ncols); err = cudaGetLastError(); }
    if (err == cudaSuccess) err = cudaMemcpyAsync(n_cols_host, d_ncols, 8, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err == cudaSuccess && *n_cols_host > 0 && *n_cols_host <= max_cols) {
        const int F = 10 * num_dtypes;
        err = cudaMemsetAsync(counts, 0, (size_t)(*n_cols_host) * F * 8, s);
        if (err == cudaSuccess) { plp_positions_kernel<<<(unsigned)((L + 255) / 256), 256, 0, s>>>(L, start, width, col_off, max_cols, major, minor); err = cudaGetLastError(); }
        if (err == cudaSuccess) {
            plp_count_kernel<<<ob, 256, 0, s>>>(n_ops, op_rec, cigar, op_ref, op_qry, op_minor, flag, mapq, dtype, seq, seq_off,
                                                min_mapq, start, end, num_dtypes, col_off,
                                                reinterpret_cast<unsigned long long *>(counts));
            err = cudaGetLastError();
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    }
    cudaFree(scratch);
    if (err != cudaSuccess) return cuda_fail(err, "pileup_counts", __FILE__, __LINE__);
    return MDK_OK;
}

}  // namespace mdk
