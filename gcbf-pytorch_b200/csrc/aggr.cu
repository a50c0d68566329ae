// K4 attention aggregation (torch_geometric AttentionalAggregation as used at reference
// gcbf/nn/gnn.py:17-19, 59-60) on a target-sorted (CSR) edge list: one warp per target node, the
// softmax statistics and the weighted sum are warp-shuffle segmented reductions over the node's
// contiguous edge range -- no atomics, deterministic.  HBM-bound: reads E*C floats once (fwd), E*C twice
// (bwd: msg for the gate gradient, d_msg written).  Plus the row gather/scatter helpers that implement
// `x[data.agent_mask]` (gcbf/algo/gcbf.py:52-53) and strided concat copies.
#include "common.cuh"

namespace gcbf {

// C = 256 channels -> each lane owns 8 channels as two float4 (lane*4 and 128 + lane*4): coalesced 512 B rows.
template <int C>
__global__ void __launch_bounds__(256) attn_aggr_fwd_kernel(const float* __restrict__ msg, int ld_msg,
                                                            const float* __restrict__ gate,
                                                            const int32_t* __restrict__ rowptr, int num_nodes,
                                                            float* __restrict__ att, float* __restrict__ aggr,
                                                            int ld_aggr) {
  static_assert(C % 128 == 0, "channels must be a multiple of 128");
  constexpr int V = C / 128;
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= num_nodes) return;
  const int e0 = rowptr[node], e1 = rowptr[node + 1];
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (e1 > e0) {
    // pass 1: max of the gate over the segment  (PyG softmax: src - max(detached src))
    float mx = -INFINITY;
    for (int e = e0 + lane; e < e1; e += 32) mx = fmaxf(mx, gate[e]);
    mx = warp_max(mx);
    // pass 2: sum of exp
    float s = 0.f;
    for (int e = e0 + lane; e < e1; e += 32) s += expf(gate[e] - mx);
    s = warp_sum(s);
    const float denom = s + 1e-16f;
    // pass 3: attention weights + weighted sum of messages
    for (int e = e0; e < e1; ++e) {
      const float a = expf(gate[e] - mx) / denom;
      if (lane == 0) att[e] = a;
      const float* row = msg + (size_t)e * ld_msg;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const float4 m = *reinterpret_cast<const float4*>(row + v * 128 + lane * 4);
        acc[v].x = fmaf(a, m.x, acc[v].x); acc[v].y = fmaf(a, m.y, acc[v].y);
        acc[v].z = fmaf(a, m.z, acc[v].z); acc[v].w = fmaf(a, m.w, acc[v].w);
      }
    }
  }
  float* out = aggr + (size_t)node * ld_aggr;
#pragma unroll
  for (int v = 0; v < V; ++v) *reinterpret_cast<float4*>(out + v * 128 + lane * 4) = acc[v];
}

// d_msg_e = att_e * d_aggr_i ;  t_e = <d_aggr_i, msg_e> ;  d_gate_e = att_e * (t_e - sum_e' att_e' t_e')
template <int C>
__global__ void __launch_bounds__(256) attn_aggr_bwd_kernel(const float* __restrict__ msg, int ld_msg,
                                                            const float* __restrict__ att,
                                                            const int32_t* __restrict__ rowptr, int num_nodes,
                                                            const float* __restrict__ d_aggr, int ld_daggr,
                                                            float* __restrict__ d_msg, int ld_dmsg,
                                                            float* __restrict__ d_gate, int accumulate) {
  constexpr int V = C / 128;
  const int node = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (node >= num_nodes) return;
  const int e0 = rowptr[node], e1 = rowptr[node + 1];
  if (e1 <= e0) return;
  float4 g[V];
  const float* grow = d_aggr + (size_t)node * ld_daggr;
#pragma unroll
  for (int v = 0; v < V; ++v) g[v] = *reinterpret_cast<const float4*>(grow + v * 128 + lane * 4);
  float wsum = 0.f;  // sum_e att_e * t_e   (identical on all lanes)
  for (int e = e0; e < e1; ++e) {
    const float a = att[e];
    const float* row = msg + (size_t)e * ld_msg;
    float* drow = d_msg + (size_t)e * ld_dmsg;
    float t = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float4 m = *reinterpret_cast<const float4*>(row + v * 128 + lane * 4);
      t = fmaf(g[v].x, m.x, t); t = fmaf(g[v].y, m.y, t); t = fmaf(g[v].z, m.z, t); t = fmaf(g[v].w, m.w, t);
      float4 d = make_float4(a * g[v].x, a * g[v].y, a * g[v].z, a * g[v].w);
      float4* dp = reinterpret_cast<float4*>(drow + v * 128 + lane * 4);
      if (accumulate) { const float4 o = *dp; d.x += o.x; d.y += o.y; d.z += o.z; d.w += o.w; }
      *dp = d;
    }
    t = warp_sum(t);
    wsum = fmaf(a, t, wsum);
    if (lane == 0) d_gate[e] = t;  // stash t_e; fixed up below
  }
  __syncwarp();
  for (int e = e0 + lane; e < e1; e += 32) d_gate[e] = att[e] * (d_gate[e] - wsum);
}

__global__ void rows_index_kernel(const float* __restrict__ src, int ld_src, const int64_t* __restrict__ idx,
                                  float* __restrict__ dst, int ld_dst, int64_t rows, int cols, int gather) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = i / cols;
  const int c = (int)(i % cols);
  if (r >= rows) return;
  const int64_t k = idx[r];
  if (gather) dst[r * ld_dst + c] = src[k * ld_src + c];
  else dst[k * ld_dst + c] = src[r * ld_src + c];
}

__global__ void copy2d_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                              int64_t rows, int cols) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = idx / cols;
  const int c = (int)(idx % cols);
  if (r >= rows) return;
  dst[r * ld_dst + c] = src[r * ld_src + c];
}

}  // namespace gcbf

using namespace gcbf;

extern "C" int gcbf_attn_aggr_fwd(const float* msg, int ld_msg, const float* gate, const int32_t* rowptr,
                                  int num_nodes, int channels, float* att, float* aggr, int ld_aggr, void* stream) {
  GCBF_REQUIRE(rowptr && aggr && num_nodes >= 0, "gcbf_attn_aggr_fwd: bad arguments");
  GCBF_REQUIRE(channels == 256, "gcbf_attn_aggr_fwd: channels=%d (only phi_dim 256 is built, gcbf/algo/gcbf.py:91)", channels);
  GCBF_REQUIRE((ld_msg & 3) == 0 && (ld_aggr & 3) == 0, "gcbf_attn_aggr_fwd: leading dims must be multiples of 4");
  if (num_nodes == 0) return GCBF_OK;
  attn_aggr_fwd_kernel<256><<<ceil_div((int64_t)num_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      msg, ld_msg, gate, rowptr, num_nodes, att, aggr, ld_aggr);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_attn_aggr_bwd(const float* msg, int ld_msg, const float* att, const int32_t* rowptr,
                                  int num_nodes, int channels, const float* d_aggr, int ld_daggr, float* d_msg,
                                  int ld_dmsg, float* d_gate, int accumulate, void* stream) {
  GCBF_REQUIRE(rowptr && d_aggr && num_nodes >= 0, "gcbf_attn_aggr_bwd: bad arguments");
  GCBF_REQUIRE(channels == 256, "gcbf_attn_aggr_bwd: channels=%d unsupported", channels);
  GCBF_REQUIRE((ld_msg & 3) == 0 && (ld_daggr & 3) == 0 && (ld_dmsg & 3) == 0, "gcbf_attn_aggr_bwd: leading dims must be multiples of 4");
  if (num_nodes == 0) return GCBF_OK;
  attn_aggr_bwd_kernel<256><<<ceil_div((int64_t)num_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      msg, ld_msg, att, rowptr, num_nodes, d_aggr, ld_daggr, d_msg, ld_dmsg, d_gate, accumulate);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_rows_gather(const float* src, int ld_src, const int64_t* idx, float* dst, int ld_dst, int64_t rows,
                                int cols, void* stream) {
  GCBF_REQUIRE(cols > 0 && rows >= 0, "gcbf_rows_gather: bad sizes");
  if (rows == 0) return GCBF_OK;
  GCBF_REQUIRE(src && dst && idx, "gcbf_rows_gather: null pointer");
  rows_index_kernel<<<ceil_div(rows * cols, 256), 256, 0, as_stream(stream)>>>(src, ld_src, idx, dst, ld_dst, rows, cols, 1);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_rows_scatter(const float* src, int ld_src, const int64_t* idx, float* dst, int ld_dst, int64_t rows,
                                 int cols, void* stream) {
  GCBF_REQUIRE(cols > 0 && rows >= 0, "gcbf_rows_scatter: bad sizes");
  if (rows == 0) return GCBF_OK;
  GCBF_REQUIRE(src && dst && idx, "gcbf_rows_scatter: null pointer");
  rows_index_kernel<<<ceil_div(rows * cols, 256), 256, 0, as_stream(stream)>>>(src, ld_src, idx, dst, ld_dst, rows, cols, 0);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_copy2d(const float* src, int ld_src, float* dst, int ld_dst, int64_t rows, int cols, void* stream) {
  GCBF_REQUIRE(cols >= 0 && rows >= 0, "gcbf_copy2d: bad sizes");
  if (rows == 0 || cols == 0) return GCBF_OK;
  GCBF_REQUIRE(src && dst, "gcbf_copy2d: null pointer");
  copy2d_kernel<<<ceil_div(rows * cols, 256), 256, 0, as_stream(stream)>>>(src, ld_src, dst, ld_dst, rows, cols);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}
