// Per-element arithmetic of the analytic h_dot pass, written once for device and host (see macbf_core.h for the
// pattern: the kernels in jvp.cu are grid-stride loops around these functions, tests/host_driver/jvp_host.cpp compiles the same
// functions with g++ for the CPU test-suite).
//
// h_dot_i = sum_k (dh_i / ds_k) . f(s_k, u_k): the directional derivative of the CBF along the closed-loop vector field with the
// graph's edges held fixed -- what the reference approximates by the finite difference (h(x + dt f) - h(x)) / dt at
// gcbf/algo/gcbf.py:193-207.  It is a forward-mode (tangent) pass through forward_graph's pieces:
//   state_dot      f(x, clamp(u + u_ref(x)))          simple_car.py:78-89, dubins_car.py:110-132, simple_drone.py:103-120
//   edge tangent   d/dt [g(s_j) - g(s_i)]             simple_car.py:246-247, dubins_car.py:724-728, simple_drone.py:313-314
//   attention      d/dt sum_e softmax(gate)_e m_e     gcbf/nn/gnn.py:17-19 (AttentionalAggregation)
// and the linear layers / activations of the MLPs, which reuse the forward GEMM kernels and gcbf_act_bwd on the tangent.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GCBF_JHD __host__ __device__ __forceinline__
#else
#define GCBF_JHD inline
#endif

namespace gcbf {
namespace jvp {

// x_dot of one node.  s: state row; uc: the node's TOTAL clamped action (agents only; ignored for obstacles); frozen: the
// single-graph reach-freeze of dynamics() (dubins_car.py:126-130, simple_drone.py:113-117).  env: 0 SimpleCar, 1 DubinsCar, 2 SimpleDrone.
GCBF_JHD void state_dot(int env, bool is_agent, const float* s, const float* uc, float speed_limit, bool frozen, float* xdot) {
  for (int k = 0; k < 6; ++k) xdot[k] = 0.f;
  if (env == 0) {
    xdot[0] = s[2]; xdot[1] = s[3]; xdot[2] = uc[0]; xdot[3] = uc[1];
  } else if (env == 1) {                               // obstacles move too (their action is zero)
    const float vc = fminf(s[3], speed_limit);
    xdot[0] = vc * cosf(s[2]);
    xdot[1] = vc * sinf(s[2]);
    if (is_agent) { xdot[2] = uc[0] * 10.f; xdot[3] = uc[1]; }
  } else if (is_agent) {                                // drone obstacles are static
    xdot[0] = s[3]; xdot[1] = s[4]; xdot[2] = s[5];
    xdot[3] = -1.1f * s[3] + 1.1f * uc[0];
    xdot[4] = -1.1f * s[4] + 1.1f * uc[1];
    xdot[5] = -6.f * s[5] + 6.f * uc[2];
  }
  if (frozen)
    for (int k = 0; k < 6; ++k) xdot[k] = 0.f;
}

// d/dt g(s) given s and s_dot.  g = identity (SimpleCar: 4, SimpleDrone: 6); DubinsCar g = [x, y, theta, v cos theta, v sin theta]
GCBF_JHD void feature_dot(int env, const float* s, const float* sd, float* gd) {
  if (env == 1) {
    const float c = cosf(s[2]), sn = sinf(s[2]);
    gd[0] = sd[0]; gd[1] = sd[1]; gd[2] = sd[2];
    gd[3] = sd[3] * c - s[3] * sn * sd[2];
    gd[4] = sd[3] * sn + s[3] * c * sd[2];
  } else {
    const int d = env == 0 ? 4 : 6;
    for (int k = 0; k < d; ++k) gd[k] = sd[k];
  }
}

// tangent of the attention aggregation of ONE (target, channel) cell over the target's CSR range [beg, end):
//   aggr = sum_e a_e m_e,  a = softmax(gate)  =>  d aggr = sum_e a_e (dm_e + m_e (dg_e - sum_k a_k dg_k))
GCBF_JHD float attn_tangent_cell(const float* msg, int ld_msg, const float* t_msg, int ld_tmsg, const float* att, const float* t_gate,
                                 int beg, int end, int c) {
  float mean_tg = 0.f;
  for (int e = beg; e < end; ++e) mean_tg += att[e] * t_gate[e];
  float acc = 0.f;
  for (int e = beg; e < end; ++e)
    acc += att[e] * (t_msg[(int64_t)e * ld_tmsg + c] + msg[(int64_t)e * ld_msg + c] * (t_gate[e] - mean_tg));
  return acc;
}

}  // namespace jvp
}  // namespace gcbf
