// fp32 CUDA-core (FFMA) implementation of the GRU layer: the --full_precision / validation path
// (MDK_PREC_FP32).  Same data flow as the tensor-core path (gru_tc.cu) with fp32 operands:
//   gi  = X . W_ih^T + folded bias           (time-parallel GEMM, gemm_fp32)
//   h_t = GRU cell(gi_t, h_{t-1} . W_hh^T)   (persistent recurrent kernel, rec_fp32)
// Reference arithmetic: torch.nn.GRU as used by medaka/architectures/gru.py:46-52,66.
#include "common.cuh"
#include "ptx.cuh"

namespace mdk {

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// -------------------------------------------------------------------------------------
// Recurrent kernel.  One CTA = NB windows of one direction; 128 threads, thread j owns hidden
// unit j: the r/z/n rows of W_hh for unit j stream from shared memory (W_hh^T resident for the
// whole sequence, 192 KiB) while the NB h vectors are shared-memory broadcasts.
// gi   [B*T][768]  (position p = b*T + t ; columns dir*384 + gate*128 + j)
// h_out[B*T][256]  (columns dir*128 + j)
// -------------------------------------------------------------------------------------
constexpr int REC_NB = 8;

__global__ void __launch_bounds__(128, 1) rec_fp32_kernel(const float *__restrict__ gi,
                                                          const float *__restrict__ w_hh_t,
                                                          const float *__restrict__ b_hn,
                                                          float *__restrict__ h_out, int64_t B, int64_t T) {
    extern __shared__ __align__(16) float smem[];
    float *wt = smem;                          // [128][384]
    float *hs = smem + H * G3;                 // [2][NB][128]
    const int j = threadIdx.x;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * REC_NB;
    const int nb = (int)min((int64_t)REC_NB, B - b0);

    const float *wsrc = w_hh_t + (int64_t)dir * H * G3;
    for (int i = j; i < H * G3 / 4; i += 128)
        reinterpret_cast<float4 *>(wt)[i] = reinterpret_cast<const float4 *>(wsrc)[i];
    for (int i = j; i < 2 * REC_NB * H; i += 128) hs[i] = 0.f;
    const float bhn = b_hn[dir * H + j];
    float hprev[REC_NB];
#pragma unroll
    for (int n = 0; n < REC_NB; ++n) hprev[n] = 0.f;
    __syncthreads();

    const int64_t col = (int64_t)dir * G3 + j;
    float gnext[3][REC_NB];
    {
        const int64_t t = dir ? (T - 1) : 0;
#pragma unroll
        for (int n = 0; n < REC_NB; ++n) {
            const bool ok = n < nb;
            const float *row = gi + ((b0 + (ok ? n : 0)) * T + t) * GI_COLS + col;
#pragma unroll
            for (int g = 0; g < 3; ++g) gnext[g][n] = ok ? ldg_stream(row + g * H) : 0.f;
        }
    }
    int cur = 0;
    for (int64_t step = 0; step < T; ++step) {
        const int64_t t = dir ? (T - 1 - step) : step;
        float gcur[3][REC_NB];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int n = 0; n < REC_NB; ++n) gcur[g][n] = gnext[g][n];
        if (step + 1 < T) {   // prefetch next step's pre-activations while this step's matvec runs
            const int64_t tn = dir ? (t - 1) : (t + 1);
#pragma unroll
            for (int n = 0; n < REC_NB; ++n) {
                const bool ok = n < nb;
                const float *row = gi + ((b0 + (ok ? n : 0)) * T + tn) * GI_COLS + col;
#pragma unroll
                for (int g = 0; g < 3; ++g) gnext[g][n] = ok ? ldg_stream(row + g * H) : 0.f;
            }
        }
        float acc[3][REC_NB];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int n = 0; n < REC_NB; ++n) acc[g][n] = 0.f;
        const float *hc = hs + cur * REC_NB * H;
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {
            float4 hv[REC_NB];
#pragma unroll
            for (int n = 0; n < REC_NB; ++n) hv[n] = *reinterpret_cast<const float4 *>(hc + n * H + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float wr = wt[(k + kk) * G3 + j];
                const float wz = wt[(k + kk) * G3 + H + j];
                const float wn = wt[(k + kk) * G3 + 2 * H + j];
#pragma unroll
                for (int n = 0; n < REC_NB; ++n) {
                    const float hvk = kk == 0 ? hv[n].x : kk == 1 ? hv[n].y : kk == 2 ? hv[n].z : hv[n].w;
                    acc[0][n] = fmaf(wr, hvk, acc[0][n]);
                    acc[1][n] = fmaf(wz, hvk, acc[1][n]);
                    acc[2][n] = fmaf(wn, hvk, acc[2][n]);
                }
            }
        }
        float *hn = hs + (cur ^ 1) * REC_NB * H;
#pragma unroll
        for (int n = 0; n < REC_NB; ++n) {
            const float r = sigmoid_acc(gcur[0][n] + acc[0][n]);
            const float z = sigmoid_acc(gcur[1][n] + acc[1][n]);
            const float nn = tanhf(gcur[2][n] + r * (acc[2][n] + bhn));
            const float h = (1.0f - z) * nn + z * hprev[n];
            hprev[n] = h;
            hn[n * H + j] = h;
            if (n < nb) h_out[((b0 + n) * T + t) * H2 + dir * H + j] = h;
        }
        __syncthreads();
        cur ^= 1;
    }
}

cudaError_t launch_rec_fp32(const float *gi, const float *w_hh_t, const float *b_hn, float *h_out, int64_t B,
                            int64_t T, cudaStream_t s) {
    if (B == 0 || T == 0) return cudaSuccess;
    const size_t smem = (size_t)(H * G3 + 2 * REC_NB * H) * sizeof(float);
    // (the attribute is per device: set it on every launch, a process may drive several GPUs from several threads)
    cudaError_t e = cudaFuncSetAttribute(rec_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((B + REC_NB - 1) / REC_NB), NDIR);
    rec_fp32_kernel<<<grid, 128, smem, s>>>(gi, w_hh_t, b_hn, h_out, B, T);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------
// Layer-1 input projection, fp32: C[P][768] = A[P][256] . W[768][256]^T + bias.
// Classic 128x128x16 shared-memory tiling, 256 threads, 8x8 register tile.
// -------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 128, GK = 16;

__global__ void __launch_bounds__(256) gemm_fp32_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                        const float *__restrict__ bias, float *__restrict__ C,
                                                        int64_t P) {
    __shared__ float As[GK][GM + 4];
    __shared__ float Ws[GK][GN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * GM;
    const int n0 = blockIdx.y * GN;
    const int tx = tid % 16, ty = tid / 16;   // 16x16 threads, each 8x8 outputs
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) acc[i][jx] = 0.f;
    // loader mapping: 128 rows x 16 k = 512 float4; thread loads 2 float4 per operand
    const int lrow = tid / 4;            // 0..63
    const int lk = (tid % 4) * 4;        // 0,4,8,12
    for (int k0 = 0; k0 < H2; k0 += GK) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int r = lrow + half * 64;
            const int64_t gm = m0 + r;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gm < P) a = *reinterpret_cast<const float4 *>(A + gm * H2 + k0 + lk);
            As[lk + 0][r] = a.x; As[lk + 1][r] = a.y; As[lk + 2][r] = a.z; As[lk + 3][r] = a.w;
            const float4 w = *reinterpret_cast<const float4 *>(W + (int64_t)(n0 + r) * H2 + k0 + lk);
            Ws[lk + 0][r] = w.x; Ws[lk + 1][r] = w.y; Ws[lk + 2][r] = w.z; Ws[lk + 3][r] = w.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[8], b[8];
            // rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}: conflict-free float4 smem reads
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Ws[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Ws[k][64 + tx * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int jx = 0; jx < 8; ++jx) acc[i][jx] = fmaf(a[i], b[jx], acc[i][jx]);
        }
        __syncthreads();
    }
    float bv[8];
#pragma unroll
    for (int jx = 0; jx < 8; ++jx) bv[jx] = bias[n0 + (jx < 4 ? tx * 4 + jx : 64 + tx * 4 + jx - 4)];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (gm >= P) continue;
        float *dst = C + gm * GI_COLS + n0;
        *reinterpret_cast<float4 *>(dst + tx * 4) = make_float4(acc[i][0] + bv[0], acc[i][1] + bv[1], acc[i][2] + bv[2], acc[i][3] + bv[3]);
        *reinterpret_cast<float4 *>(dst + 64 + tx * 4) = make_float4(acc[i][4] + bv[4], acc[i][5] + bv[5], acc[i][6] + bv[6], acc[i][7] + bv[7]);
    }
}

cudaError_t launch_gemm_fp32(const float *A, const float *W, const float *bias, float *C, int64_t P,
                             cudaStream_t s) {
    if (P == 0) return cudaSuccess;
    dim3 grid((unsigned)((P + GM - 1) / GM), GI_COLS / GN);
    gemm_fp32_kernel<<<grid, 256, 0, s>>>(A, W, bias, C, P);
    return cudaGetLastError();
}

}  // namespace mdk
