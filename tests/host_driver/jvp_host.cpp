// TEST INFRASTRUCTURE ONLY: host build of the per-element functions of the analytic h_dot kernels (gcbf-pytorch_b200/csrc/jvp_core.h);
// each function is the serial form of the corresponding kernel's grid-stride loop in csrc/jvp.cu.  Compiled by tests/test_jvp_cpu.py.
#include <math.h>
#include <stdint.h>
#include "jvp_core.h"

using namespace gcbf::jvp;

extern "C" {

void host_state_dot(int env, int num_graphs, int N, int n, const float* states, int ld, const float* action, const float* u_ref, const float* goal,
                    int ld_goal, int goal_gstride, float action_lim, float speed_limit, float dist2goal, int freeze, float* out, int ld_out) {
  const int sd = env == 2 ? 6 : 4, ad = env == 2 ? 3 : 2, pd = env == 2 ? 3 : 2;
  for (int64_t node = 0; node < (int64_t)num_graphs * N; ++node) {
    const int g = (int)(node / N), l = (int)(node % N);
    const bool is_agent = l < n;
    float s[6] = {0, 0, 0, 0, 0, 0}, uc[3] = {0, 0, 0}, xd[6];
    for (int k = 0; k < sd; ++k) s[k] = states[node * ld + k];
    bool frozen = false;
    if (is_agent) {
      const int64_t a = (int64_t)g * n + l;
      for (int k = 0; k < ad; ++k) {
        const float raw = action[a * ad + k] + u_ref[a * ad + k];
        uc[k] = fminf(fmaxf(raw, -action_lim), action_lim);
      }
      if (freeze && env != 0) {
        float acc = 0.f;
        for (int k = 0; k < pd; ++k) {
          const float d = s[k] - goal[((int64_t)g * goal_gstride + l) * ld_goal + k];
          acc = fmaf(d, d, acc);
        }
        frozen = sqrtf(acc) < dist2goal;
      }
    }
    state_dot(env, is_agent, s, uc, speed_limit, frozen, xd);
    for (int k = 0; k < sd; ++k) out[node * ld_out + k] = xd[k];
  }
}

void host_edge_attr_tangent(int env, const float* states, int ld, const float* sdot, int ld_sd, const int64_t* ei, int64_t E, float* out) {
  const int sd = env == 2 ? 6 : 4, ed = env == 0 ? 4 : (env == 1 ? 5 : 6);
  for (int64_t e = 0; e < E; ++e) {
    const int64_t j = ei[e], i = ei[E + e];
    float sj[6], dj[6], si[6], di[6], gj[6], gi[6];
    for (int k = 0; k < sd; ++k) { sj[k] = states[j * ld + k]; dj[k] = sdot[j * ld_sd + k]; si[k] = states[i * ld + k]; di[k] = sdot[i * ld_sd + k]; }
    feature_dot(env, sj, dj, gj);
    feature_dot(env, si, di, gi);
    for (int k = 0; k < ed; ++k) out[e * ed + k] = gj[k] - gi[k];
  }
}

void host_attn_aggr_tangent(const float* msg, int ld_msg, const float* t_msg, int ld_tmsg, const float* att, const float* t_gate, const int32_t* rowptr,
                            int num_nodes, int C, float* out, int ld_out) {
  for (int i = 0; i < num_nodes; ++i)
    for (int c = 0; c < C; ++c) out[(int64_t)i * ld_out + c] = attn_tangent_cell(msg, ld_msg, t_msg, ld_tmsg, att, t_gate, rowptr[i], rowptr[i + 1], c);
}

}  // extern "C"
