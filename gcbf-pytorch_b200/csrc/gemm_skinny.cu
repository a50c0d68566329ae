// Skinny linear layers: in-features K <= 16 -- the first phi layer, cat[x_i, x_j, e_ij] (12..14 wide) -> 2048
// (reference gcbf/nn/gnn.py:31, gcbf/nn/mlp.py:44-47).  These are HBM-bound streams over the [M, 2048] side
// (E x 2048 x 4 B = 198 MB at config C2) with 2*K flops per element, so they get dedicated kernels instead of a
// 128x128 GEMM tile that would be 90 % padding:
//   fwd   : Y[M,N]  = act(alpha * X[M,K] W[N,K]^T + b)       thread = output column, X rows broadcast from smem
//   dgrad : dX[M,K] = alpha * dZ[M,N] W[N,K]                 warp = row, lanes sweep N, shuffle-reduce K sums
//   wgrad : dW[N,K] += alpha * dZ^T X ; db[N] += colsum(dZ)  thread = column n, rows split over blockIdx.y
#include "common.cuh"

namespace gcbf {

constexpr int SK = 16;          // max in-features handled here
constexpr int SK_ROWS = 64;     // rows staged per block iteration

__global__ void __launch_bounds__(256) skinny_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                         int ldw, const float* __restrict__ bias,
                                                         const float* __restrict__ alpha_p, float* __restrict__ Y, int ldy,
                                                         int M, int N, int K, int act, int rows_per_block) {
  __shared__ __align__(16) float xs[SK_ROWS][SK];
  const int n = blockIdx.x * 256 + threadIdx.x;
  const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
  float w[SK];
#pragma unroll
  for (int k = 0; k < SK; ++k) w[k] = (n < N && k < K) ? __ldg(W + (size_t)n * ldw + k) * alpha : 0.f;
  const float b = (n < N && bias) ? __ldg(bias + n) : 0.f;
  const int m_begin = blockIdx.y * rows_per_block, m_end = min(M, m_begin + rows_per_block);
  for (int m0 = m_begin; m0 < m_end; m0 += SK_ROWS) {
    __syncthreads();
    for (int i = threadIdx.x; i < SK_ROWS * SK; i += 256) {
      const int r = i / SK, k = i % SK;
      xs[r][k] = (m0 + r < m_end && k < K) ? X[(size_t)(m0 + r) * ldx + k] : 0.f;
    }
    __syncthreads();
    if (n < N) {
      const int rows = min(SK_ROWS, m_end - m0);
      for (int r = 0; r < rows; ++r) {
        const float4* xr = reinterpret_cast<const float4*>(xs[r]);
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < SK / 4; ++q) {
          const float4 v = xr[q];
          acc = fmaf(v.x, w[4 * q], acc); acc = fmaf(v.y, w[4 * q + 1], acc);
          acc = fmaf(v.z, w[4 * q + 2], acc); acc = fmaf(v.w, w[4 * q + 3], acc);
        }
        float y = acc + b;
        if (act == GCBF_ACT_RELU) y = fmaxf(y, 0.f);
        else if (act == GCBF_ACT_TANH) y = tanhf(y);
        Y[(size_t)(m0 + r) * ldy + n] = y;
      }
    }
  }
}

__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ dZ, int lddz, const float* __restrict__ W,
                                                           int ldw, const float* __restrict__ alpha_p,
                                                           const float* __restrict__ relu_src, int ld_relu,
                                                           float* __restrict__ dX, int lddx, int M, int N, int K, int accumulate) {
  const int m = (blockIdx.x * 256 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float acc[SK];
#pragma unroll
  for (int k = 0; k < SK; ++k) acc[k] = 0.f;
  const float* zrow = dZ + (size_t)m * lddz;
  for (int n = lane; n < N; n += 32) {
    const float z = zrow[n];
    const float* wr = W + (size_t)n * ldw;
#pragma unroll
    for (int k = 0; k < SK; ++k)
      if (k < K) acc[k] = fmaf(z, __ldg(wr + k), acc[k]);
  }
  const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
#pragma unroll
  for (int k = 0; k < SK; ++k) {
    const float s = warp_sum(acc[k]);
    if (lane == k && k < K) {
      float v = alpha * s;
      if (relu_src && !(relu_src[(size_t)m * ld_relu + k] > 0.f)) v = 0.f;
      float* dst = dX + (size_t)m * lddx + k;
      *dst = accumulate ? *dst + v : v;
    }
  }
}

__global__ void __launch_bounds__(256) skinny_wgrad_kernel(const float* __restrict__ dZ, int lddz, const float* __restrict__ X,
                                                           int ldx, const float* __restrict__ alpha_p, float* __restrict__ dW,
                                                           int lddw, float* __restrict__ db, int M, int N, int K,
                                                           int rows_per_block) {
  __shared__ __align__(16) float xs[SK_ROWS][SK];
  const int n = blockIdx.x * 256 + threadIdx.x;
  float acc[SK];
#pragma unroll
  for (int k = 0; k < SK; ++k) acc[k] = 0.f;
  float accb = 0.f;
  const int m_begin = blockIdx.y * rows_per_block, m_end = min(M, m_begin + rows_per_block);
  for (int m0 = m_begin; m0 < m_end; m0 += SK_ROWS) {
    __syncthreads();
    for (int i = threadIdx.x; i < SK_ROWS * SK; i += 256) {
      const int r = i / SK, k = i % SK;
      xs[r][k] = (m0 + r < m_end && k < K) ? X[(size_t)(m0 + r) * ldx + k] : 0.f;
    }
    __syncthreads();
    if (n < N) {
      const int rows = min(SK_ROWS, m_end - m0);
      for (int r = 0; r < rows; ++r) {
        const float z = dZ[(size_t)(m0 + r) * lddz + n];
        const float4* xr = reinterpret_cast<const float4*>(xs[r]);
        accb += z;
#pragma unroll
        for (int q = 0; q < SK / 4; ++q) {
          const float4 v = xr[q];
          acc[4 * q] = fmaf(z, v.x, acc[4 * q]); acc[4 * q + 1] = fmaf(z, v.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(z, v.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(z, v.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (n < N) {
    const float alpha = alpha_p ? __ldg(alpha_p) : 1.f;
#pragma unroll
    for (int k = 0; k < SK; ++k)
      if (k < K) atomicAdd(dW + (size_t)n * lddw + k, alpha * acc[k]);
    if (db) atomicAdd(db + n, accb);
  }
}

bool skinny_supported(int K) { return K <= SK; }

int launch_skinny_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* inv_sigma, float* Y,
                      int ldy, int M, int N, int K, int act, cudaStream_t st) {
  const int col_blocks = ceil_div(N, 256);
  int row_blocks = max(1, min(ceil_div(M, SK_ROWS), (4 * kNumSMs) / col_blocks));
  const int rpb = ceil_div(ceil_div(M, row_blocks), SK_ROWS) * SK_ROWS;
  row_blocks = ceil_div(M, rpb);
  skinny_fwd_kernel<<<dim3(col_blocks, row_blocks), 256, 0, st>>>(X, ldx, W, ldw, bias, inv_sigma, Y, ldy, M, N, K, act, rpb);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

int launch_skinny_dgrad(const float* dZ, int lddz, const float* W, int ldw, const float* inv_sigma, const float* relu_src,
                        int ld_relu, float* dX, int lddx, int M, int N, int K, int accumulate, cudaStream_t st) {
  skinny_dgrad_kernel<<<ceil_div((int64_t)M * 32, 256), 256, 0, st>>>(dZ, lddz, W, ldw, inv_sigma, relu_src, ld_relu, dX, lddx,
                                                                     M, N, K, accumulate);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

int launch_skinny_wgrad(const float* dZ, int lddz, const float* X, int ldx, const float* inv_sigma, float* dW, int lddw,
                        float* db, int M, int N, int K, int accumulate, cudaStream_t st) {
  if (!accumulate) {
    GCBF_CUDA_OK(cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)K * 4, N, st));
    if (db) GCBF_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * 4, st));
  }
  const int col_blocks = ceil_div(N, 256);
  int row_blocks = max(1, min(ceil_div(M, 4 * SK_ROWS), (4 * kNumSMs) / col_blocks));
  const int rpb = ceil_div(ceil_div(M, row_blocks), SK_ROWS) * SK_ROWS;
  row_blocks = ceil_div(M, rpb);
  skinny_wgrad_kernel<<<dim3(col_blocks, row_blocks), 256, 0, st>>>(dZ, lddz, X, ldx, inv_sigma, dW, lddw, db, M, N, K, rpb);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

}  // namespace gcbf
