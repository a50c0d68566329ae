// K7 spectral-norm power iteration (+ the gradient term through sigma) and K8 clip_grad_norm_ + Adam on a
// flat fp32 bucket.  All of these are HBM-bound streaming kernels: the power iteration reads W twice
// (W^T u, then W v), the fix-up reads dW and W once and rewrites dW, clip+Adam reads p,g,m,v and writes
// p,m,v (28 B/param, the algorithmic minimum).
#include "common.cuh"

namespace gcbf {

constexpr int kSnRowSplit = 16;

// partial[by][k] = sum_{rows of slice by} W[r][k] * u[r]
__global__ void sn_wt_u_kernel(const float* __restrict__ W, int ldw, int N, int K, const float* __restrict__ u,
                               float* __restrict__ partial, int rows_per_block) {
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float s = 0.f;
  if (col < K)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s = fmaf(W[(size_t)r * ldw + col], u[r], s);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < K) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    partial[(size_t)blockIdx.y * K + col] = t;
  }
}

__device__ __forceinline__ float block_sum(float v, float* sm) {  // blockDim.x == 1024
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < 32) ? sm[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) sm[32] = t;
  __syncthreads();
  t = sm[32];
  __syncthreads();
  return t;
}

// v = normalize(sum_slices partial), eps 1e-12 (F.normalize: x / max(||x||, eps))
__global__ void sn_finalize_v_kernel(const float* __restrict__ partial, int nsplit, int K, float* __restrict__ v) {
  __shared__ float sm[33];
  float sq = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < nsplit; ++s) t += partial[(size_t)s * K + k];
    v[k] = t;
    sq = fmaf(t, t, sq);
  }
  const float nrm = fmaxf(sqrtf(block_sum(sq, sm)), 1e-12f);
  for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = v[k] / nrm;
}

// s[n] = W[n, :] . v   (one warp per row)
__global__ void sn_w_v_kernel(const float* __restrict__ W, int ldw, int N, int K, const float* __restrict__ v,
                              float* __restrict__ s) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(W[(size_t)row * ldw + k], v[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) s[row] = acc;
}

// u = normalize(s); sigma = u . s; inv_sigma = 1 / sigma
__global__ void sn_finalize_u_kernel(const float* __restrict__ s, int N, float* __restrict__ u,
                                     float* __restrict__ inv_sigma) {
  __shared__ float sm[33];
  float sq = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) sq = fmaf(s[n], s[n], sq);
  const float nrm = fmaxf(sqrtf(block_sum(sq, sm)), 1e-12f);
  float dot = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float un = s[n] / nrm;
    u[n] = un;
    dot = fmaf(un, s[n], dot);
  }
  const float sigma = block_sum(dot, sm);
  if (threadIdx.x == 0) *inv_sigma = 1.f / sigma;
}

// ---- batched power iteration: every spectral-normalised layer of a net in four launches (blockIdx.z = layer) -------------
constexpr int kSnMaxBatch = 16;
struct SnBatch {
  const float* W[kSnMaxBatch];
  float* u[kSnMaxBatch];
  float* v[kSnMaxBatch];
  float* inv_sigma[kSnMaxBatch];
  float* partial[kSnMaxBatch];
  float* s[kSnMaxBatch];
  int ldw[kSnMaxBatch], N[kSnMaxBatch], K[kSnMaxBatch], rpb[kSnMaxBatch], nsplit[kSnMaxBatch];
};

__global__ void sn_wt_u_batched_kernel(const __grid_constant__ SnBatch b) {
  const int l = blockIdx.z;
  const int N = b.N[l], K = b.K[l], ldw = b.ldw[l], rows_per_block = b.rpb[l];
  if ((int)blockIdx.y >= b.nsplit[l] || (int)blockIdx.x * 32 >= K) return;
  const float* __restrict__ W = b.W[l];
  const float* __restrict__ u = b.u[l];
  __shared__ float sm[8][33];
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float s = 0.f;
  if (col < K)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s = fmaf(W[(size_t)r * ldw + col], u[r], s);
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < K) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
    b.partial[l][(size_t)blockIdx.y * K + col] = t;
  }
}

__global__ void sn_finalize_v_batched_kernel(const __grid_constant__ SnBatch b) {
  const int l = blockIdx.x;
  const int K = b.K[l], nsplit = b.nsplit[l];
  const float* __restrict__ partial = b.partial[l];
  float* __restrict__ v = b.v[l];
  __shared__ float sm[33];
  float sq = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < nsplit; ++s) t += partial[(size_t)s * K + k];
    v[k] = t;
    sq = fmaf(t, t, sq);
  }
  const float nrm = fmaxf(sqrtf(block_sum(sq, sm)), 1e-12f);
  for (int k = threadIdx.x; k < K; k += blockDim.x) v[k] = v[k] / nrm;
}

__global__ void sn_w_v_batched_kernel(const __grid_constant__ SnBatch b) {
  const int l = blockIdx.y;
  const int N = b.N[l], K = b.K[l], ldw = b.ldw[l];
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* __restrict__ W = b.W[l];
  const float* __restrict__ v = b.v[l];
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(W[(size_t)row * ldw + k], v[k], acc);
  acc = warp_sum(acc);
  if (lane == 0) b.s[l][row] = acc;
}

__global__ void sn_finalize_u_batched_kernel(const __grid_constant__ SnBatch b) {
  const int l = blockIdx.x;
  const int N = b.N[l];
  const float* __restrict__ s = b.s[l];
  float* __restrict__ u = b.u[l];
  __shared__ float sm[33];
  float sq = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) sq = fmaf(s[n], s[n], sq);
  const float nrm = fmaxf(sqrtf(block_sum(sq, sm)), 1e-12f);
  float dot = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float un = s[n] / nrm;
    u[n] = un;
    dot = fmaf(un, s[n], dot);
  }
  const float sigma = block_sum(dot, sm);
  if (threadIdx.x == 0) *b.inv_sigma[l] = 1.f / sigma;
}

__global__ void dot_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, int N, int K,
                           double* __restrict__ out) {
  double acc = 0;
  const int64_t total = (int64_t)N * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / K, c = i % K;
    acc += (double)A[r * lda + c] * (double)B[r * ldb + c];
  }
  __shared__ double sm[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(out, t);
  }
}

// dW_orig = dW - <dW, W> * inv_sigma * u v^T (the gradient through sigma of torch spectral_norm).  acc == nullptr: in
// place on dW; otherwise the corrected gradient is ADDED to acc (the parameter's .grad view) and dW is left alone.
__global__ void sn_rank1_kernel(float* __restrict__ dW, int lddw, int N, int K, const float* __restrict__ u,
                                const float* __restrict__ v, const float* __restrict__ inv_sigma,
                                const double* __restrict__ inner, float* __restrict__ acc, int ldacc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * K) return;
  const int64_t r = i / K, c = i % K;
  const float coef = (float)(*inner) * (*inv_sigma);
  const float g = dW[r * lddw + c] - coef * u[r] * v[c];
  if (acc) acc[r * ldacc + c] += g;
  else dW[r * lddw + c] = g;
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t count, double* __restrict__ out) {
  double acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const double x = g[i];
    acc += x * x;
  }
  __shared__ double sm[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    atomicAdd(out, t);
  }
}

// torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (_multi_tensor_adam op order), reference
// gcbf/algo/gcbf.py:223-226
__global__ void clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, int64_t count, const double* __restrict__ sumsq,
                                 float max_norm, float one_minus_b1, float b2, float one_minus_b2, float bc2_sqrt,
                                 float eps, float neg_step_size) {
  const float total = (float)sqrt(*sumsq);
  const float coef = fminf(max_norm / (total + 1e-6f), 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = __fmul_rn(g[i], coef);
    const float mi = fmaf(one_minus_b1, __fsub_rn(gi, m[i]), m[i]);                  // exp_avg.lerp_(g, 1-b1)
    const float vi = __fadd_rn(__fmul_rn(v[i], b2), __fmul_rn(__fmul_rn(one_minus_b2, gi), gi));  // mul_, addcmul_
    const float denom = __fadd_rn(__fsqrt_rn(vi) / bc2_sqrt, eps);
    m[i] = mi;
    v[i] = vi;
    p[i] = __fadd_rn(p[i], __fmul_rn(neg_step_size, mi / denom));                    // addcdiv_
  }
}

}  // namespace gcbf

using namespace gcbf;

extern "C" size_t gcbf_sn_workspace_floats(int N, int K) {
  return (size_t)kSnRowSplit * (size_t)K + (size_t)N + 16;
}

extern "C" int gcbf_sn_power_iter(const float* W, int ldw, int N, int K, float* u, float* v, float* inv_sigma,
                                  float* workspace, void* stream) {
  GCBF_REQUIRE(W && u && v && inv_sigma && workspace && N > 0 && K > 0 && ldw >= K, "gcbf_sn_power_iter: bad arguments");
  cudaStream_t st = as_stream(stream);
  int rows_per_block = ceil_div(N, kSnRowSplit);
  const int nsplit = ceil_div(N, rows_per_block);
  float* partial = workspace;
  float* s = workspace + (size_t)kSnRowSplit * K;
  dim3 g1(ceil_div(K, 32), nsplit), b1(32, 8);
  sn_wt_u_kernel<<<g1, b1, 0, st>>>(W, ldw, N, K, u, partial, rows_per_block);
  GCBF_LAUNCH_OK();
  sn_finalize_v_kernel<<<1, 1024, 0, st>>>(partial, nsplit, K, v);
  GCBF_LAUNCH_OK();
  sn_w_v_kernel<<<ceil_div((int64_t)N * 32, 256), 256, 0, st>>>(W, ldw, N, K, v, s);
  GCBF_LAUNCH_OK();
  sn_finalize_u_kernel<<<1, 1024, 0, st>>>(s, N, u, inv_sigma);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

// Same arithmetic as gcbf_sn_power_iter per layer (bit-identical u, v, 1/sigma), `count` layers per call.  `layers` is a
// HOST array; workspace_floats >= sum over layers of gcbf_sn_workspace_floats(N, K).
extern "C" int gcbf_sn_power_iter_batched(const gcbf_sn_layer* layers, int count, float* workspace, size_t workspace_floats,
                                          void* stream) {
  GCBF_REQUIRE(layers && count >= 0 && workspace, "gcbf_sn_power_iter_batched: bad arguments");
  cudaStream_t st = as_stream(stream);
  size_t off = 0;
  for (int base = 0; base < count; base += kSnMaxBatch) {
    const int nb = min(kSnMaxBatch, count - base);
    SnBatch b{};
    int max_ct = 0, max_rb = 0;
    for (int i = 0; i < nb; ++i) {
      const gcbf_sn_layer& L = layers[base + i];
      GCBF_REQUIRE(L.W && L.u && L.v && L.inv_sigma && L.N > 0 && L.K > 0 && L.ldw >= L.K, "gcbf_sn_power_iter_batched: layer %d", base + i);
      const size_t need = gcbf_sn_workspace_floats(L.N, L.K);
      GCBF_REQUIRE(off + need <= workspace_floats, "gcbf_sn_power_iter_batched: workspace too small");
      b.W[i] = L.W; b.u[i] = L.u; b.v[i] = L.v; b.inv_sigma[i] = L.inv_sigma; b.ldw[i] = L.ldw; b.N[i] = L.N; b.K[i] = L.K;
      b.rpb[i] = ceil_div(L.N, kSnRowSplit);
      b.nsplit[i] = ceil_div(L.N, b.rpb[i]);
      b.partial[i] = workspace + off;
      b.s[i] = workspace + off + (size_t)kSnRowSplit * L.K;
      off += need;
      max_ct = max(max_ct, ceil_div(L.K, 32));
      max_rb = max(max_rb, ceil_div((int64_t)L.N * 32, 256));
    }
    sn_wt_u_batched_kernel<<<dim3(max_ct, kSnRowSplit, nb), dim3(32, 8), 0, st>>>(b);
    GCBF_LAUNCH_OK();
    sn_finalize_v_batched_kernel<<<nb, 1024, 0, st>>>(b);
    GCBF_LAUNCH_OK();
    sn_w_v_batched_kernel<<<dim3(max_rb, nb), 256, 0, st>>>(b);
    GCBF_LAUNCH_OK();
    sn_finalize_u_batched_kernel<<<nb, 1024, 0, st>>>(b);
    GCBF_LAUNCH_OK();
  }
  return GCBF_OK;
}

extern "C" int gcbf_sn_grad_fixup(float* dW, int lddw, const float* W, int ldw, int N, int K, const float* u,
                                  const float* v, const float* inv_sigma, float* workspace, float* acc, int ldacc,
                                  void* stream) {
  GCBF_REQUIRE(dW && W && u && v && inv_sigma && workspace && N > 0 && K > 0, "gcbf_sn_grad_fixup: bad arguments");
  GCBF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7) == 0, "gcbf_sn_grad_fixup: workspace must be 8-byte aligned");
  cudaStream_t st = as_stream(stream);
  double* inner = reinterpret_cast<double*>(workspace);
  GCBF_CUDA_OK(cudaMemsetAsync(inner, 0, sizeof(double), st));
  const int64_t total = (int64_t)N * K;
  dot_kernel<<<(int)imin64(ceil_div(total, 256), 4 * kNumSMs), 256, 0, st>>>(dW, lddw, W, ldw, N, K, inner);
  GCBF_LAUNCH_OK();
  GCBF_REQUIRE(!acc || ldacc >= K, "gcbf_sn_grad_fixup: ldacc");
  sn_rank1_kernel<<<ceil_div(total, 256), 256, 0, st>>>(dW, lddw, N, K, u, v, inv_sigma, inner, acc, ldacc);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_grad_sumsq(const float* g, int64_t count, double* sumsq, void* stream) {
  GCBF_REQUIRE(sumsq && count >= 0 && (g || count == 0), "gcbf_grad_sumsq: bad arguments");
  if (count == 0) return GCBF_OK;
  sumsq_kernel<<<(int)imin64(ceil_div(count, 256), 8 * kNumSMs), 256, 0, as_stream(stream)>>>(g, count, sumsq);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_clip_adam(float* p, const float* g, float* m, float* v, int64_t count, const double* sumsq,
                              double max_norm, double lr, double beta1, double beta2, double eps, int step,
                              void* stream) {
  GCBF_REQUIRE(count >= 0 && step >= 1 && sumsq, "gcbf_clip_adam: bad arguments");
  if (count == 0) return GCBF_OK;
  GCBF_REQUIRE(p && g && m && v, "gcbf_clip_adam: null pointer");
  // python-double scalar arithmetic of torch/optim/adam.py, then cast to fp32 where torch hands it to a kernel
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const double step_size = lr / bc1, bc2_sqrt = sqrt(bc2);
  clip_adam_kernel<<<(int)imin64(ceil_div(count, 256), 16 * kNumSMs), 256, 0, as_stream(stream)>>>(
      p, g, m, v, count, sumsq, (float)max_norm, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
      (float)bc2_sqrt, (float)eps, (float)(-step_size));
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}
