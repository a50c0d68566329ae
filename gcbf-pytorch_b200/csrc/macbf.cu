// MACBF path (reference gcbf/algo/macbf.py:20-239, gcbf/nn/gnn.py:82-135): the kernels the baseline algorithm needs
// beyond the GCBF ones -- top-k filtered radius graph, per-edge safe / unsafe masks, max aggregation (forward + argmax-routed
// backward) and the per-edge losses.  The small MLPs (widths 64 / 128) run on the same linear kernels as the ends of the GCBF MLPs.
//
// Everything here is HBM / L2 streaming with trivial arithmetic; the kernels are grid-stride loops around the per-element
// functions of macbf_core.h, which the CPU test-suite compiles for the host and checks against the reference.
#include "common.cuh"
#include "macbf_kernels.cuh"

namespace gcbf {

using namespace macbf;

constexpr int kNP = 11;   // partial sums in use (MLP_SUM_UNSAFE .. MLP_CNT_AGENTS)

__global__ void macbf_loss_partials_kernel(const float* __restrict__ h, const float* __restrict__ hn, const uint8_t* __restrict__ safe,
                                           const uint8_t* __restrict__ unsafe, int64_t E, const float* __restrict__ act, int ad,
                                           int64_t M, float alpha, float eps, float dt, double* __restrict__ partial) {
  double acc[kNP];
#pragma unroll
  for (int k = 0; k < kNP; ++k) acc[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t e = first; e < E; e += stride) edge_terms(h[e], hn[e], safe[e], unsafe[e], alpha, eps, dt, acc);
  for (int64_t i = first; i < M; i += stride) {
    acc[MLP_SUM_ACT] += action_term(act + i * ad, ad);
    acc[MLP_CNT_AGENTS] += 1.0;
  }
  __shared__ double sm[kNP][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kNP; ++k) {
    const double v = warp_sum(acc[k]);
    if (lane == 0) sm[k][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < kNP) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[threadIdx.x][w];
    atomicAdd(partial + threadIdx.x, t);
  }
}

}  // namespace gcbf

using namespace gcbf;

static int check_graph_args(const char* who, const float* states, int ld_state, int pos_dim, int num_graphs, int nodes_per_graph,
                            int num_agents, int metric, int max_neighbors) {
  GCBF_REQUIRE(states != nullptr, "%s: null pointer", who);
  GCBF_REQUIRE(pos_dim >= 1 && pos_dim <= 3 && ld_state >= pos_dim, "%s: pos_dim=%d ld=%d", who, pos_dim, ld_state);
  GCBF_REQUIRE(num_graphs >= 0 && nodes_per_graph >= num_agents && num_agents >= 0, "%s: bad sizes", who);
  GCBF_REQUIRE(metric == 0 || metric == 1, "%s: metric %d", who, metric);
  GCBF_REQUIRE(max_neighbors >= 1, "%s: max_neighbors %d", who, max_neighbors);
  GCBF_REQUIRE((int64_t)num_graphs * num_agents < (1ll << 31), "%s: too many agents", who);
  return GCBF_OK;
}

extern "C" int gcbf_radius_graph_topk_count(const float* states, int ld_state, int pos_dim, int num_graphs, int nodes_per_graph,
                                            int num_agents, float radius, int metric, int max_neighbors, int32_t* rowptr, void* stream) {
  GCBF_REQUIRE(rowptr != nullptr, "gcbf_radius_graph_topk_count: null rowptr");
  if (int rc = check_graph_args("gcbf_radius_graph_topk_count", states, ld_state, pos_dim, num_graphs, nodes_per_graph, num_agents, metric,
                                max_neighbors)) return rc;
  cudaStream_t st = as_stream(stream);
  const int64_t na = (int64_t)num_graphs * num_agents;
  if (na > 0) {
    radius_topk_kernel<false><<<ceil_div(na, 128), 128, 0, st>>>(states, ld_state, pos_dim, num_graphs, nodes_per_graph, num_agents,
                                                                 radius, metric, max_neighbors, rowptr, nullptr, nullptr, 0);
    GCBF_LAUNCH_OK();
  }
  GCBF_CUDA_OK(exclusive_scan_i32(rowptr, (int)na, st));
  return GCBF_OK;
}

extern "C" int gcbf_radius_graph_topk_fill(const float* states, int ld_state, int pos_dim, int num_graphs, int nodes_per_graph,
                                           int num_agents, float radius, int metric, int max_neighbors, const int32_t* rowptr,
                                           int64_t* edge_index, int64_t num_edges, void* stream) {
  GCBF_REQUIRE(rowptr && (edge_index || num_edges == 0) && num_edges >= 0, "gcbf_radius_graph_topk_fill: null pointer");
  if (int rc = check_graph_args("gcbf_radius_graph_topk_fill", states, ld_state, pos_dim, num_graphs, nodes_per_graph, num_agents, metric,
                                max_neighbors)) return rc;
  const int64_t na = (int64_t)num_graphs * num_agents;
  if (na == 0 || num_edges == 0) return GCBF_OK;
  radius_topk_kernel<true><<<ceil_div(na, 128), 128, 0, as_stream(stream)>>>(states, ld_state, pos_dim, num_graphs, nodes_per_graph,
                                                                            num_agents, radius, metric, max_neighbors, nullptr, rowptr,
                                                                            edge_index, num_edges);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_edge_masks(const float* edge_attr, int ld_edge_attr, int pos_dim, int64_t num_edges, double agent_radius,
                               uint8_t* safe, uint8_t* unsafe, void* stream) {
  GCBF_REQUIRE(num_edges >= 0 && pos_dim >= 1 && pos_dim <= 3 && ld_edge_attr >= pos_dim, "gcbf_edge_masks: bad sizes");
  if (num_edges == 0) return GCBF_OK;
  GCBF_REQUIRE(edge_attr && safe && unsafe, "gcbf_edge_masks: null pointer");
  // thresholds are python doubles cast to fp32 by torch's scalar comparison: 4R (safe), 2R (collision)
  const float safe_thr = (float)(4 * agent_radius), coll_thr = (float)(2 * agent_radius);
  const int grid = (int)imin64(ceil_div(num_edges, 256), 8 * kNumSMs);
  edge_masks_kernel<<<grid, 256, 0, as_stream(stream)>>>(edge_attr, ld_edge_attr, pos_dim, num_edges, safe_thr, coll_thr, safe, unsafe);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_seg_max_fwd(const float* msg, int ld_msg, const int32_t* rowptr, int num_nodes, int channels, float* out, int ld_out,
                                int32_t* argmax, void* stream) {
  GCBF_REQUIRE(num_nodes >= 0 && channels >= 1 && ld_msg >= channels && ld_out >= channels, "gcbf_seg_max_fwd: bad sizes");
  if (num_nodes == 0) return GCBF_OK;
  GCBF_REQUIRE(rowptr && out && argmax, "gcbf_seg_max_fwd: null pointer");   // msg may be null when the graph has no edges
  const int64_t total = (int64_t)num_nodes * channels;
  const int grid = (int)imin64(ceil_div(total, 256), 8 * kNumSMs);
  seg_max_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(msg, ld_msg, rowptr, num_nodes, channels, out, ld_out, argmax);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_seg_max_bwd(const float* d_out, int ld_dout, const int32_t* argmax, int num_nodes, int channels, float* d_msg,
                                int ld_dmsg, int64_t num_edges, void* stream) {
  GCBF_REQUIRE(num_nodes >= 0 && channels >= 1 && ld_dout >= channels && ld_dmsg >= channels && num_edges >= 0, "gcbf_seg_max_bwd: bad sizes");
  if (num_edges == 0) return GCBF_OK;
  GCBF_REQUIRE(d_out && argmax && d_msg, "gcbf_seg_max_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  GCBF_CUDA_OK(cudaMemsetAsync(d_msg, 0, (size_t)num_edges * ld_dmsg * sizeof(float), st));
  if (num_nodes == 0) return GCBF_OK;
  const int64_t total = (int64_t)num_nodes * channels;
  const int grid = (int)imin64(ceil_div(total, 256), 8 * kNumSMs);
  seg_max_bwd_kernel<<<grid, 256, 0, st>>>(d_out, ld_dout, argmax, num_nodes, channels, d_msg, ld_dmsg);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_macbf_loss_partials(const float* h, const float* h_next, const uint8_t* safe, const uint8_t* unsafe, int64_t num_edges,
                                        const float* action, int action_dim, int64_t num_agents, float alpha, float eps, float dt,
                                        double* partial, void* stream) {
  GCBF_REQUIRE(partial && num_edges >= 0 && num_agents >= 0 && action_dim >= 0 && action_dim <= 8, "gcbf_macbf_loss_partials: bad arguments");
  cudaStream_t st = as_stream(stream);
  GCBF_CUDA_OK(cudaMemsetAsync(partial, 0, MLP_SIZE * sizeof(double), st));
  if (num_edges == 0 && num_agents == 0) return GCBF_OK;
  GCBF_REQUIRE(num_edges == 0 || (h && h_next && safe && unsafe), "gcbf_macbf_loss_partials: null edge pointer");
  GCBF_REQUIRE(num_agents == 0 || action, "gcbf_macbf_loss_partials: null action pointer");
  const int grid = (int)imin64(ceil_div(imax64(num_edges, num_agents), 256), 4 * kNumSMs);
  macbf_loss_partials_kernel<<<grid, 256, 0, st>>>(h, h_next, safe, unsafe, num_edges, action, action_dim, num_agents, alpha, eps, dt, partial);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_macbf_loss_grads(const float* h, const float* h_next, const uint8_t* safe, const uint8_t* unsafe, int64_t num_edges,
                                     const float* action, int action_dim, int64_t num_agents, float alpha, float eps, float dt,
                                     float coef_unsafe, float coef_safe, float coef_hdot, float coef_action, const double* partial,
                                     float* d_h, float* d_h_next, float* d_action, float* scalars, void* stream) {
  GCBF_REQUIRE(partial && scalars && num_edges >= 0 && num_agents >= 0 && action_dim >= 0, "gcbf_macbf_loss_grads: bad arguments");
  GCBF_REQUIRE(num_edges == 0 || (h && h_next && safe && unsafe && d_h && d_h_next), "gcbf_macbf_loss_grads: null edge pointer");
  GCBF_REQUIRE(num_agents == 0 || (action && d_action), "gcbf_macbf_loss_grads: null action pointer");
  const int64_t work = imax64(imax64(num_edges, num_agents * action_dim), 1);
  const int grid = (int)imin64(ceil_div(work, 256), 8 * kNumSMs);
  macbf_loss_grads_kernel<<<grid, 256, 0, as_stream(stream)>>>(h, h_next, safe, unsafe, num_edges, action, action_dim, num_agents, alpha,
                                                               eps, dt, coef_unsafe, coef_safe, coef_hdot, coef_action, partial, d_h,
                                                               d_h_next, d_action, scalars);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}
