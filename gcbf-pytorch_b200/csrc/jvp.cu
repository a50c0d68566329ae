// Analytic h_dot: the three kernels a tangent pass through forward_graph + CBFGNN needs beyond the forward GEMMs --
// closed-loop state derivative, edge-feature tangent, attention-aggregation tangent.  Grid-stride loops around jvp_core.h.
#include "common.cuh"
#include "jvp_core.h"

namespace gcbf {

__global__ void state_dot_kernel(int env, int num_graphs, int N, int n, const float* __restrict__ states, int ld,
                                 const float* __restrict__ action, const float* __restrict__ u_ref, const float* __restrict__ goal, int ld_goal,
                                 int goal_gstride, float action_lim, float speed_limit, float dist2goal, int freeze,
                                 float* __restrict__ out, int ld_out) {
  const int sd = env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4, ad = env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2, pd = env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
  const int64_t total = (int64_t)num_graphs * N;
  for (int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; node < total; node += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(node / N), l = (int)(node % N);
    const bool is_agent = l < n;
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, uc[3] = {0.f, 0.f, 0.f}, xd[6];
    for (int k = 0; k < sd; ++k) s[k] = states[node * ld + k];
    bool frozen = false;
    if (is_agent) {
      const int64_t a = (int64_t)g * n + l;
      for (int k = 0; k < ad; ++k) {
        const float raw = __fadd_rn(action[a * ad + k], u_ref[a * ad + k]);
        uc[k] = fminf(fmaxf(raw, -action_lim), action_lim);
      }
      if (freeze && env != GCBF_ENV_SIMPLE_CAR) {
        float acc = 0.f;
        for (int k = 0; k < pd; ++k) {
          const float d = __fsub_rn(s[k], goal[((int64_t)g * goal_gstride + l) * ld_goal + k]);
          acc = __fmaf_rn(d, d, acc);
        }
        frozen = __fsqrt_rn(acc) < dist2goal;
      }
    }
    jvp::state_dot(env, is_agent, s, uc, speed_limit, frozen, xd);
    for (int k = 0; k < sd; ++k) out[node * ld_out + k] = xd[k];
  }
}

__global__ void edge_attr_tangent_kernel(int env, const float* __restrict__ states, int ld, const float* __restrict__ sdot, int ld_sd,
                                         const int64_t* __restrict__ ei, int64_t E, float* __restrict__ out) {
  const int sd = env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4, ed = env == GCBF_ENV_SIMPLE_CAR ? 4 : (env == GCBF_ENV_DUBINS_CAR ? 5 : 6);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = ei[e], i = ei[E + e];                // source j, target i: edge_attr = g(s_j) - g(s_i)
    float sj[6], dj[6], si[6], di[6], gj[6], gi[6];
    for (int k = 0; k < sd; ++k) {
      sj[k] = states[j * ld + k]; dj[k] = sdot[j * ld_sd + k];
      si[k] = states[i * ld + k]; di[k] = sdot[i * ld_sd + k];
    }
    jvp::feature_dot(env, sj, dj, gj);
    jvp::feature_dot(env, si, di, gi);
    for (int k = 0; k < ed; ++k) out[e * ed + k] = gj[k] - gi[k];
  }
}

// thread per (target, channel), channel fastest
__global__ void attn_tangent_kernel(const float* __restrict__ msg, int ld_msg, const float* __restrict__ t_msg, int ld_tmsg,
                                    const float* __restrict__ att, const float* __restrict__ t_gate, const int32_t* __restrict__ rowptr,
                                    int num_nodes, int C, float* __restrict__ out, int ld_out) {
  const int64_t total = (int64_t)num_nodes * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / C), c = (int)(idx % C);
    out[(int64_t)i * ld_out + c] = jvp::attn_tangent_cell(msg, ld_msg, t_msg, ld_tmsg, att, t_gate, rowptr[i], rowptr[i + 1], c);
  }
}

}  // namespace gcbf

using namespace gcbf;

extern "C" int gcbf_state_dot(const gcbf_env_cfg* cfg, const float* states, int ld_state, const float* action, const float* u_ref,
                              const float* goal, int ld_goal, int goal_per_graph, int freeze, float* state_dot, int ld_out, void* stream) {
  GCBF_REQUIRE(cfg != nullptr, "gcbf_state_dot: null cfg");
  GCBF_REQUIRE(cfg->env >= 0 && cfg->env <= 2, "gcbf_state_dot: unknown env %d", cfg->env);
  GCBF_REQUIRE(cfg->num_graphs >= 0 && cfg->num_agents >= 0 && cfg->nodes_per_graph >= cfg->num_agents, "gcbf_state_dot: bad sizes");
  const int sd = cfg->env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4, pd = cfg->env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
  GCBF_REQUIRE(ld_state >= sd && ld_out >= sd, "gcbf_state_dot: leading dimensions");
  const int64_t nodes = (int64_t)cfg->num_graphs * cfg->nodes_per_graph;
  GCBF_REQUIRE(nodes < (1ll << 31), "gcbf_state_dot: too many nodes");
  if (nodes == 0) return GCBF_OK;
  GCBF_REQUIRE(states && state_dot && (cfg->num_agents == 0 || (action && u_ref)), "gcbf_state_dot: null pointer");
  GCBF_REQUIRE(!freeze || cfg->env == GCBF_ENV_SIMPLE_CAR || (goal && ld_goal >= pd), "gcbf_state_dot: the reach-freeze needs the goal positions");
  const float action_lim = cfg->env == GCBF_ENV_DUBINS_CAR ? 2.f : 10.f;       // simple_car.py:264-268, dubins_car.py:758-762, simple_drone.py:343-347
  const int grid = (int)imin64(ceil_div(nodes, 256), 8 * kNumSMs);
  state_dot_kernel<<<grid, 256, 0, as_stream(stream)>>>(cfg->env, cfg->num_graphs, cfg->nodes_per_graph, cfg->num_agents, states, ld_state, action,
                                                        u_ref, goal, ld_goal, goal_per_graph ? cfg->num_agents : 0, action_lim,
                                                        (float)cfg->speed_limit, (float)cfg->dist2goal, freeze, state_dot, ld_out);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_edge_attr_tangent(int env, const float* states, int ld_state, const float* state_dot, int ld_sdot, const int64_t* edge_index,
                                      int64_t num_edges, float* t_edge_attr, void* stream) {
  GCBF_REQUIRE(env >= 0 && env <= 2 && num_edges >= 0, "gcbf_edge_attr_tangent: bad arguments");
  const int sd = env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4;
  GCBF_REQUIRE(ld_state >= sd && ld_sdot >= sd, "gcbf_edge_attr_tangent: leading dimensions");
  if (num_edges == 0) return GCBF_OK;
  GCBF_REQUIRE(states && state_dot && edge_index && t_edge_attr, "gcbf_edge_attr_tangent: null pointer");
  const int grid = (int)imin64(ceil_div(num_edges, 256), 8 * kNumSMs);
  edge_attr_tangent_kernel<<<grid, 256, 0, as_stream(stream)>>>(env, states, ld_state, state_dot, ld_sdot, edge_index, num_edges, t_edge_attr);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_attn_aggr_tangent(const float* msg, int ld_msg, const float* t_msg, int ld_tmsg, const float* att, const float* t_gate,
                                      const int32_t* rowptr, int num_nodes, int channels, float* t_aggr, int ld_taggr, void* stream) {
  GCBF_REQUIRE(num_nodes >= 0 && channels >= 1 && ld_msg >= channels && ld_tmsg >= channels && ld_taggr >= channels, "gcbf_attn_aggr_tangent: bad sizes");
  if (num_nodes == 0) return GCBF_OK;
  GCBF_REQUIRE(rowptr && t_aggr, "gcbf_attn_aggr_tangent: null pointer");     // the edge arrays may be null for a graph without edges
  const int64_t total = (int64_t)num_nodes * channels;
  const int grid = (int)imin64(ceil_div(total, 256), 8 * kNumSMs);
  attn_tangent_kernel<<<grid, 256, 0, as_stream(stream)>>>(msg, ld_msg, t_msg, ld_tmsg, att, t_gate, rowptr, num_nodes, channels, t_aggr, ld_taggr);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}
