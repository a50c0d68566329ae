// Kernels of the analytic h_dot pass: grid-stride loops around the per-element functions of jvp_core.h.  Kept in a header
// of their own, free of CUDA runtime includes, so that tests/host_driver/jvp_grid.cpp can compile the SAME kernel bodies for the host
// on an emulated grid (tests/host_driver/cuda_emu.h) -- the build container has no GPU.
#pragma once
#include "gcbf_b200.h"
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
