// Linear layers with FEW ROWS (M <= 64): the rollout-time forward of a single graph (reference gcbf/algo/gcbf.py:128-139: 16 agents,
// ~45 edges per env step) and the plumbing config C1.  With so few rows a layer is a stream over its WEIGHTS (2048 x 2048 fp32 = 16 MB,
// resident in the 126 MB L2 from one pass to the next), not a GEMM tile problem: a 128 x 128 tile grid puts 16 CTAs on 148 SMs.
//   fwd   : Y[M,N]   = act(alpha * X W^T + b)      warp = 2 output columns, lanes split K (16-byte loads of W rows), shuffle-reduce
//   dgrad : dX[M,K] (+)= alpha * dZ W (* mask)      thread = 4 consecutive k (16-byte loads of W rows), N split over blockIdx.y,
//                                                   partial sums (already masked / scaled: both are linear) added atomically
//   wgrad : dW[N,K] (+)= alpha * dZ^T X             thread = (n, 4 consecutive k): a streaming write of dW, X / dZ tiles in smem
// Exact fp32 FFMA arithmetic like gemm_simt.cu (different summation order).
#include "common.cuh"

namespace gcbf {

constexpr int FR_MAX_M = 64;
constexpr int FR_MT = 16;           // rows per register tile

bool fewrows_supported(int M, int N, int K) { return M >= 1 && M <= FR_MAX_M && N >= 64 && K >= 32; }

// ---- forward -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fewrows_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, int ldw,
                                                          const float* __restrict__ bias, const float* __restrict__ alpha_p,
                                                          float* __restrict__ Y, int ldy, int M, int N, int K, int act, int vec_ok) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (blockIdx.x * 8 + warp) * 2;
  if (n0 >= N) return;
  const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
  const bool two = n0 + 1 < N;
  const float* w0 = W + (size_t)n0 * ldw;
  const float* w1 = W + (size_t)(two ? n0 + 1 : n0) * ldw;
  for (int m0 = 0; m0 < M; m0 += FR_MT) {
    float acc0[FR_MT], acc1[FR_MT];
#pragma unroll
    for (int r = 0; r < FR_MT; ++r) { acc0[r] = 0.f; acc1[r] = 0.f; }
    if (vec_ok) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w0 + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(w1 + k));
#pragma unroll
        for (int r = 0; r < FR_MT; ++r) {
          if (m0 + r < M) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * ldx + k));
            acc0[r] = fmaf(x.x, a.x, acc0[r]); acc0[r] = fmaf(x.y, a.y, acc0[r]); acc0[r] = fmaf(x.z, a.z, acc0[r]); acc0[r] = fmaf(x.w, a.w, acc0[r]);
            acc1[r] = fmaf(x.x, b.x, acc1[r]); acc1[r] = fmaf(x.y, b.y, acc1[r]); acc1[r] = fmaf(x.z, b.z, acc1[r]); acc1[r] = fmaf(x.w, b.w, acc1[r]);
          }
        }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float a = __ldg(w0 + k), b = __ldg(w1 + k);
#pragma unroll
        for (int r = 0; r < FR_MT; ++r) {
          if (m0 + r < M) {
            const float x = __ldg(X + (size_t)(m0 + r) * ldx + k);
            acc0[r] = fmaf(x, a, acc0[r]);
            acc1[r] = fmaf(x, b, acc1[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < FR_MT; ++r) { acc0[r] = warp_sum(acc0[r]); acc1[r] = warp_sum(acc1[r]); }
    // lane r writes row m0 + r (both columns)
#pragma unroll
    for (int r = 0; r < FR_MT; ++r) {
      if (lane == r && m0 + r < M) {
        float y0 = alpha * acc0[r] + (bias ? __ldg(bias + n0) : 0.f);
        float y1 = alpha * acc1[r] + ((bias && two) ? __ldg(bias + n0 + 1) : 0.f);
        if (act == GCBF_ACT_RELU) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
        else if (act == GCBF_ACT_TANH) { y0 = tanhf(y0); y1 = tanhf(y1); }
        float* d = Y + (size_t)(m0 + r) * ldy + n0;
        d[0] = y0;
        if (two) d[1] = y1;
      }
    }
  }
}

// ---- data gradient -------------------------------------------------------------------------------------------------------------
constexpr int FRD_NSLICE = 32;      // rows of W (contraction indices) per block

__global__ void __launch_bounds__(128) fewrows_dgrad_kernel(const float* __restrict__ dZ, int lddz, const float* __restrict__ W, int ldw,
                                                            const float* __restrict__ alpha_p, const float* __restrict__ relu_src,
                                                            int ld_relu, float* __restrict__ dX, int lddx, int M, int N, int K, int m0) {
  __shared__ float dzs[FR_MT][FRD_NSLICE];
  const int k = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int nb = blockIdx.y * FRD_NSLICE, ne = min(N, nb + FRD_NSLICE);
  for (int i = threadIdx.x; i < FR_MT * FRD_NSLICE; i += 128) {
    const int r = i / FRD_NSLICE, n = i % FRD_NSLICE;
    dzs[r][n] = (m0 + r < M && nb + n < N) ? dZ[(size_t)(m0 + r) * lddz + nb + n] : 0.f;
  }
  __syncthreads();
  if (k >= K) return;
  const bool full = (k + 3 < K) && ((ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  float acc[FR_MT][4];
#pragma unroll
  for (int r = 0; r < FR_MT; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f; }
  for (int n = nb; n < ne; ++n) {
    float4 w;
    const float* wr = W + (size_t)n * ldw + k;
    if (full) w = __ldg(reinterpret_cast<const float4*>(wr));
    else { w.x = wr[0]; w.y = (k + 1 < K) ? wr[1] : 0.f; w.z = (k + 2 < K) ? wr[2] : 0.f; w.w = (k + 3 < K) ? wr[3] : 0.f; }
#pragma unroll
    for (int r = 0; r < FR_MT; ++r) {
      const float z = dzs[r][n - nb];
      acc[r][0] = fmaf(z, w.x, acc[r][0]); acc[r][1] = fmaf(z, w.y, acc[r][1]);
      acc[r][2] = fmaf(z, w.z, acc[r][2]); acc[r][3] = fmaf(z, w.w, acc[r][3]);
    }
  }
  const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
#pragma unroll
  for (int r = 0; r < FR_MT; ++r) {
    if (m0 + r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (k + j >= K) continue;
      float v = alpha * acc[r][j];
      if (relu_src && !(relu_src[(size_t)(m0 + r) * ld_relu + k + j] > 0.f)) v = 0.f;      // mask and alpha are linear: applied per partial
      if (v != 0.f) atomicAdd(dX + (size_t)(m0 + r) * lddx + k + j, v);
    }
  }
}

// ---- weight gradient -----------------------------------------------------------------------------------------------------------
constexpr int FRW_KT = 128, FRW_NT = 32;

__global__ void __launch_bounds__(256) fewrows_wgrad_kernel(const float* __restrict__ dZ, int lddz, const float* __restrict__ X, int ldx,
                                                            const float* __restrict__ alpha_p, float* __restrict__ dW, int lddw, int M,
                                                            int N, int K, int accumulate) {
  __shared__ __align__(16) float xs[FR_MAX_M][FRW_KT];
  __shared__ float zs[FR_MAX_M][FRW_NT + 1];
  const int k0 = blockIdx.x * FRW_KT, n0 = blockIdx.y * FRW_NT;
  for (int i = threadIdx.x; i < M * FRW_KT; i += 256) {
    const int r = i / FRW_KT, c = i % FRW_KT;
    xs[r][c] = (k0 + c < K) ? X[(size_t)r * ldx + k0 + c] : 0.f;
  }
  for (int i = threadIdx.x; i < M * FRW_NT; i += 256) {
    const int r = i / FRW_NT, c = i % FRW_NT;
    zs[r][c] = (n0 + c < N) ? dZ[(size_t)r * lddz + n0 + c] : 0.f;
  }
  __syncthreads();
  const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
  const int kc = (threadIdx.x & 31) * 4;
  const bool vec = ((lddw & 3) == 0) && ((reinterpret_cast<uintptr_t>(dW) & 15) == 0) && (k0 + kc + 3 < K);
#pragma unroll
  for (int pass = 0; pass < FRW_NT / 8; ++pass) {
    const int nl = pass * 8 + (threadIdx.x >> 5);
    const int n = n0 + nl;
    if (n >= N || k0 + kc >= K) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < M; ++m) {
      const float z = zs[m][nl];
      const float4 x = *reinterpret_cast<const float4*>(&xs[m][kc]);
      acc.x = fmaf(z, x.x, acc.x); acc.y = fmaf(z, x.y, acc.y); acc.z = fmaf(z, x.z, acc.z); acc.w = fmaf(z, x.w, acc.w);
    }
    acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha;
    float* d = dW + (size_t)n * lddw + k0 + kc;
    if (vec) {
      if (accumulate) { const float4 o = *reinterpret_cast<const float4*>(d); acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
      *reinterpret_cast<float4*>(d) = acc;
    } else {
      const float v[4] = {acc.x, acc.y, acc.z, acc.w};
      for (int j = 0; j < 4; ++j)
        if (k0 + kc + j < K) d[j] = accumulate ? d[j] + v[j] : v[j];
    }
  }
}

int launch_colsum(const float* dZ, int lddz, int M, int N, float* db, int accumulate, cudaStream_t st);

int launch_fewrows_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* inv_sigma, float* Y, int ldy, int M,
                       int N, int K, int act, cudaStream_t st) {
  const int vec_ok = ((K & 3) == 0 && (ldx & 3) == 0 && (ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(W) & 15) == 0) ? 1 : 0;
  fewrows_fwd_kernel<<<ceil_div(N, 16), 256, 0, st>>>(X, ldx, W, ldw, bias, inv_sigma, Y, ldy, M, N, K, act, vec_ok);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

int launch_fewrows_dgrad(const float* dZ, int lddz, const float* W, int ldw, const float* inv_sigma, const float* relu_src, int ld_relu,
                         float* dX, int lddx, int M, int N, int K, int accumulate, cudaStream_t st) {
  if (!accumulate) GCBF_CUDA_OK(cudaMemset2DAsync(dX, (size_t)lddx * 4, 0, (size_t)K * 4, M, st));
  for (int m0 = 0; m0 < M; m0 += FR_MT) {
    fewrows_dgrad_kernel<<<dim3(ceil_div(K, 512), ceil_div(N, FRD_NSLICE)), 128, 0, st>>>(dZ, lddz, W, ldw, inv_sigma, relu_src, ld_relu, dX, lddx, M,
                                                                                       N, K, m0);
    GCBF_LAUNCH_OK();
  }
  return GCBF_OK;
}

int launch_fewrows_wgrad(const float* dZ, int lddz, const float* X, int ldx, const float* inv_sigma, float* dW, int lddw, float* db, int M, int N,
                         int K, int accumulate, cudaStream_t st) {
  fewrows_wgrad_kernel<<<dim3(ceil_div(K, FRW_KT), ceil_div(N, FRW_NT)), 256, 0, st>>>(dZ, lddz, X, ldx, inv_sigma, dW, lddw, M, N, K, accumulate);
  GCBF_LAUNCH_OK();
  if (db) return launch_colsum(dZ, lddz, M, N, db, accumulate, st);
  return GCBF_OK;
}

}  // namespace gcbf
