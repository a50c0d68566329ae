// Analytic h_dot: the three kernels a tangent pass through forward_graph + CBFGNN needs beyond the forward GEMMs --
// closed-loop state derivative, edge-feature tangent, attention-aggregation tangent.  The kernels live in jvp_kernels.cuh (shared with the
// host emulation of the CPU test-suite); this file holds the C-ABI entry points that launch them.
#include "common.cuh"
#include "jvp_kernels.cuh"



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
