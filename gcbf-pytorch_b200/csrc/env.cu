// K5 (nominal controller + one finite-difference dynamics step, forward and VJP to the action) and the
// safe/unsafe masks of K6, for the three reference environments.  All arithmetic is elementwise fp32 in
// the reference's operation order; multiplications / additions the reference performs as separate ATen ops
// are kept unfused (__fmul_rn / __fadd_rn) so that x+ -- which feeds the bit-exact radius graph -- does not
// pick up FMA contraction differences.
#include "common.cuh"

namespace gcbf {

struct EnvCfg {
  int env, num_graphs, N, n;
  float speed_limit, dist2goal, action_lim, dt;
  float safe_thr, coll_thr, warn_thr, diag_safe, diag_unsafe, two_r;
};

__device__ __forceinline__ float norm_fma(const float* d, int k) {  // torch.norm on CPU: fma chain + sqrt
  float acc = 0.f;
  for (int i = 0; i < k; ++i) acc = __fmaf_rn(d[i], d[i], acc);
  return __fsqrt_rn(acc);
}

__device__ __forceinline__ float torch_remainder(float a, float b) {  // c10: fmod, then fix the sign
  float m = fmodf(a, b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) m += b;
  return m;
}

// ---- nominal controller ------------------------------------------------------------------------
// s: state of one agent; goal row; K: LQR gain [a, s] (row-major) or nullptr.  out: u_ref[a].
__device__ __forceinline__ void u_ref_one(const EnvCfg& c, const float* s, const float* goal, const float* K,
                                          float* out) {
  if (c.env == GCBF_ENV_SIMPLE_CAR) {
    // reference gcbf/env/simple_car.py:270-304
    const float diff[4] = {__fsub_rn(s[0], goal[0]), __fsub_rn(s[1], goal[1]), s[2], s[3]};  // goal velocity = 0
    for (int u = 0; u < 2; ++u) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc = fmaf(K[u * 4 + k], diff[k], acc);
      out[u] = -acc;
    }
    const float vel[2] = {s[2], s[3]};
    const float vn = norm_fma(vel, 2);
    if (__fsub_rn(vn, c.speed_limit) > 0.f) {
      const float over = __fsub_rn(vn, c.speed_limit);
      for (int u = 0; u < 2; ++u) out[u] = __fsub_rn(out[u], __fmul_rn(__fmul_rn(over, vel[u] / vn), 50.f));
    }
  } else if (c.env == GCBF_ENV_SIMPLE_DRONE) {
    // reference gcbf/env/simple_drone.py:349-377
    float diff[6];
    for (int k = 0; k < 6; ++k) diff[k] = __fsub_rn(s[k], goal[k]);
    for (int u = 0; u < 3; ++u) {
      float acc = 0.f;
      for (int k = 0; k < 6; ++k) acc = fmaf(K[u * 6 + k], diff[k], acc);
      out[u] = -acc;
    }
    const float vel[3] = {s[3], s[4], s[5]};
    const float vn = norm_fma(vel, 3);
    if (__fsub_rn(vn, c.speed_limit) > 0.f) {
      const float over = __fsub_rn(vn, c.speed_limit);
      for (int u = 0; u < 3; ++u) out[u] = __fsub_rn(out[u], __fmul_rn(__fmul_rn(over, vel[u] / vn), 10.f));
    }
  } else {
    // reference gcbf/env/dubins_car.py:764-816 (PID to the goal)
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    const float k_omega = 0.2f, k_v = 0.3f, k_a = 0.6f;
    const float d0 = __fsub_rn(s[0], goal[0]), d1 = __fsub_rn(s[1], goal[1]);
    const float dxy[2] = {d0, d1};
    const float dist = norm_fma(dxy, 2);
    const float den = __fadd_rn(dist, 0.0001f);
    const float sgn = (-d1 > 0.f) ? 1.f : ((-d1 < 0.f) ? -1.f : 0.f);
    const float theta_t = torch_remainder(__fmul_rn(acosf(-d0 / den), sgn), two_pi);
    const float theta = torch_remainder(s[2], two_pi);
    const float theta_diff = __fsub_rn(theta_t, theta);
    const float ct = cosf(theta), st = sinf(theta);
    float dot = __fadd_rn(__fmul_rn(-d0, ct), __fmul_rn(-d1, st));
    float cosb = fminf(fmaxf(dot / den, -1.f), 1.f);
    const float tb = acosf(cosb);
    float omega;
    if (theta <= pi) {
      const bool in_a = (theta_diff < pi) && (theta_diff >= 0.f);
      omega = in_a ? __fmul_rn(k_omega, tb) : __fmul_rn(-k_omega, tb);
    } else {
      const bool in_b = (theta_diff > -pi) && (theta_diff <= 0.f);
      omega = in_b ? __fmul_rn(-k_omega, tb) : __fmul_rn(k_omega, tb);
    }
    omega = fminf(fmaxf(omega, -5.f), 5.f);
    float a = __fadd_rn(__fmul_rn(-k_a, s[3]), __fmul_rn(k_v, dist));
    if (__fsub_rn(s[3], c.speed_limit) > 0.f) a = fminf(a, 0.f);
    if (__fadd_rn(s[3], c.speed_limit) < 0.f) a = fmaxf(a, 0.f);
    out[0] = omega;
    out[1] = a;
  }
}

__global__ void u_ref_kernel(EnvCfg c, const float* __restrict__ states, int ld, const float* __restrict__ goal,
                             int ld_goal, int goal_gstride, const float* __restrict__ K, float* __restrict__ out) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= c.num_graphs * c.n) return;
  const int g = a / c.n, il = a % c.n;
  const int sd = c.env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4;
  const int ad = c.env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
  float s[6], gl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, u[3];   // goal rows narrower than 6 read as zeros
  for (int k = 0; k < sd; ++k) s[k] = states[((size_t)g * c.N + il) * ld + k];
  for (int k = 0; k < ld_goal && k < 6; ++k) gl[k] = goal[((size_t)g * goal_gstride + il) * ld_goal + k];   // goal_gstride 0: one goal set for all graphs
  u_ref_one(c, s, gl, K, u);
  for (int k = 0; k < ad; ++k) out[(size_t)a * ad + k] = u[k];
}

// ---- one dynamics step: x+ = x + dt * f(x, clamp(u + u_ref(x))) ------------------------------------
__global__ void step_fwd_kernel(EnvCfg c, const float* __restrict__ states, int ld, const float* __restrict__ action,
                                const float* __restrict__ goal, int ld_goal, int goal_gstride, const float* __restrict__ K, int freeze,
                                float* __restrict__ next, uint8_t* __restrict__ pass_mask) {
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= (int64_t)c.num_graphs * c.N) return;
  const int g = (int)(node / c.N), l = (int)(node % c.N);
  const bool is_agent = l < c.n;
  const int sd = c.env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4;
  const int ad = c.env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
  float s[6], xdot[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, uc[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < sd; ++k) s[k] = states[node * ld + k];
  bool frozen = false;
  if (is_agent) {
    const int a = g * c.n + l;
    float gl[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ur[3];
    for (int k = 0; k < ld_goal && k < 6; ++k) gl[k] = goal[((size_t)g * goal_gstride + l) * ld_goal + k];
    u_ref_one(c, s, gl, K, ur);
    for (int k = 0; k < ad; ++k) {
      const float raw = __fadd_rn(action[(size_t)a * ad + k], ur[k]);
      uc[k] = fminf(fmaxf(raw, -c.action_lim), c.action_lim);
      // torch.clamp backward passes the gradient where min <= x <= max (NaN -> no gradient)
      pass_mask[(size_t)a * ad + k] = (raw >= -c.action_lim && raw <= c.action_lim) ? 1 : 0;
    }
    if (freeze && c.env != GCBF_ENV_SIMPLE_CAR) {
      // single-graph branch of dynamics(): dubins_car.py:126-130, simple_drone.py:113-117
      const int pd = c.env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
      float d[3];
      for (int k = 0; k < pd; ++k) d[k] = __fsub_rn(s[k], gl[k]);
      frozen = norm_fma(d, pd) < c.dist2goal;
      if (frozen)
        for (int k = 0; k < ad; ++k) pass_mask[(size_t)a * ad + k] = 0;
    }
  }
  if (c.env == GCBF_ENV_SIMPLE_CAR) {            // simple_car.py:78-89
    xdot[0] = s[2]; xdot[1] = s[3]; xdot[2] = uc[0]; xdot[3] = uc[1];
  } else if (c.env == GCBF_ENV_DUBINS_CAR) {     // dubins_car.py:110-132 (obstacles move too)
    const float vc = fminf(s[3], c.speed_limit);
    xdot[0] = __fmul_rn(vc, cosf(s[2]));
    xdot[1] = __fmul_rn(vc, sinf(s[2]));
    if (is_agent) { xdot[2] = __fmul_rn(uc[0], 10.f); xdot[3] = uc[1]; }
  } else {                                        // simple_drone.py:103-120 (obstacles are static)
    if (is_agent) {
      xdot[0] = s[3]; xdot[1] = s[4]; xdot[2] = s[5];
      xdot[3] = __fadd_rn(__fmul_rn(-1.1f, s[3]), __fmul_rn(1.1f, uc[0]));
      xdot[4] = __fadd_rn(__fmul_rn(-1.1f, s[4]), __fmul_rn(1.1f, uc[1]));
      xdot[5] = __fadd_rn(__fmul_rn(-6.f, s[5]), __fmul_rn(6.f, uc[2]));
    }
  }
  for (int k = 0; k < sd; ++k) {
    const float xd = frozen ? __fmul_rn(xdot[k], 0.f) : xdot[k];
    next[node * ld + k] = __fadd_rn(s[k], __fmul_rn(xd, c.dt));   // gcbf/env/base.py:397-398
  }
}

__global__ void step_bwd_kernel(EnvCfg c, const float* __restrict__ d_next, int ld,
                                const uint8_t* __restrict__ pass_mask, float* __restrict__ d_action) {
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= c.num_graphs * c.n) return;
  const int g = a / c.n, l = a % c.n;
  const float* d = d_next + ((size_t)g * c.N + l) * ld;
  if (c.env == GCBF_ENV_SIMPLE_CAR) {
    for (int k = 0; k < 2; ++k) d_action[(size_t)a * 2 + k] = pass_mask[(size_t)a * 2 + k] ? d[2 + k] * c.dt : 0.f;
  } else if (c.env == GCBF_ENV_DUBINS_CAR) {
    d_action[(size_t)a * 2 + 0] = pass_mask[(size_t)a * 2 + 0] ? (d[2] * c.dt) * 10.f : 0.f;
    d_action[(size_t)a * 2 + 1] = pass_mask[(size_t)a * 2 + 1] ? d[3] * c.dt : 0.f;
  } else {
    const float b[3] = {1.1f, 1.1f, 6.f};
    for (int k = 0; k < 3; ++k)
      d_action[(size_t)a * 3 + k] = pass_mask[(size_t)a * 3 + k] ? (d[3 + k] * c.dt) * b[k] : 0.f;
  }
}

// ---- safe / unsafe masks: one warp per agent sweeps the nodes of its graph --------------------------
__global__ void masks_kernel(EnvCfg c, const float* __restrict__ states, int ld, uint8_t* __restrict__ safe,
                             uint8_t* __restrict__ unsafe, uint8_t* __restrict__ collision_out) {
  const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= c.num_graphs * c.n) return;
  const int g = a / c.n, il = a % c.n;
  const int sd = c.env == GCBF_ENV_SIMPLE_DRONE ? 6 : 4;
  const int pd = c.env == GCBF_ENV_SIMPLE_DRONE ? 3 : 2;
  const float* base = states + (size_t)g * c.N * ld;
  float si[6];
  for (int k = 0; k < sd; ++k) si[k] = base[(size_t)il * ld + k];
  float tv[3] = {0.f, 0.f, 0.f};
  if (c.env == GCBF_ENV_SIMPLE_CAR) {               // simple_car.py:354-358
    const float vel[2] = {si[2], si[3]};
    const float v = __fadd_rn(norm_fma(vel, 2), 0.00001f);
    tv[0] = si[2] / v; tv[1] = si[3] / v;
  } else if (c.env == GCBF_ENV_DUBINS_CAR) {        // dubins_car.py:866-868
    tv[0] = cosf(si[2]); tv[1] = sinf(si[2]);
  } else {                                          // simple_drone.py:429-435 (z component NOT normalised)
    const float vel[3] = {si[3], si[4], si[5]};
    const float v = __fadd_rn(norm_fma(vel, 3), 0.00001f);
    tv[0] = si[3] / v; tv[1] = si[4] / v; tv[2] = si[5];
  }
  bool all_safe = true, any_unsafe = false, any_coll = false;
  for (int j = lane; j < c.N; j += 32) {
    float d[3];
    for (int k = 0; k < pd; ++k) d[k] = __fsub_rn(si[k], base[(size_t)j * ld + k]);   // pos_i - pos_j
    const float nrm = norm_fma(d, pd);
    const float dist_s = (j == il) ? __fadd_rn(nrm, c.diag_safe) : nrm;     // safe_mask diagonal
    const float dist_u = (j == il) ? __fadd_rn(nrm, c.diag_unsafe) : nrm;   // unsafe_mask diagonal
    all_safe = all_safe && (dist_s > c.safe_thr);
    const bool collision = dist_u < c.coll_thr;
    const bool warn = dist_u < c.warn_thr;
    const float den = __fadd_rn(nrm, 0.0001f);
    float inner = 0.f;
    for (int k = 0; k < pd; ++k) inner = __fadd_rn(inner, __fmul_rn(-(d[k] / den), tv[k]));
    const float thr = cosf(asinf(c.two_r / __fadd_rn(dist_u, 0.0000001f)));
    any_unsafe = any_unsafe || collision || ((inner > thr) && warn);
    any_coll = any_coll || collision;
  }
  all_safe = __all_sync(0xffffffffu, all_safe);
  any_unsafe = __any_sync(0xffffffffu, any_unsafe);
  any_coll = __any_sync(0xffffffffu, any_coll);
  if (lane == 0) {
    safe[a] = all_safe ? 1 : 0;
    unsafe[a] = any_unsafe ? 1 : 0;
    if (collision_out) collision_out[a] = any_coll ? 1 : 0;
  }
}

static int make_cfg(const gcbf_env_cfg* p, EnvCfg* c, const char* who) {
  GCBF_REQUIRE(p != nullptr, "%s: null cfg", who);
  GCBF_REQUIRE(p->env >= 0 && p->env <= 2, "%s: unknown env %d", who, p->env);
  GCBF_REQUIRE(p->num_graphs >= 0 && p->num_agents >= 0 && p->nodes_per_graph >= p->num_agents, "%s: bad sizes", who);
  GCBF_REQUIRE((int64_t)p->num_graphs * p->nodes_per_graph < (1ll << 31), "%s: too many nodes", who);
  c->env = p->env; c->num_graphs = p->num_graphs; c->N = p->nodes_per_graph; c->n = p->num_agents;
  c->speed_limit = (float)p->speed_limit; c->dist2goal = (float)p->dist2goal; c->dt = (float)p->dt;
  const double R = p->agent_radius;
  // per-env constants exactly as hard-coded in the reference (python doubles, then cast to fp32 by torch):
  //   action_lim: simple_car.py:264-268 (10), dubins_car.py:758-762 (2), simple_drone.py:343-347 (10)
  //   safe: > 4R / 3R / 4R (simple_car.py:325, dubins_car.py:837, simple_drone.py:398), diagonal += 4R+1
  //   unsafe: collision < 2R; warn 4R / 3R / 4R; diagonal += 4R+1 / 4R+1 / 2R+1 (simple_drone.py:426)
  c->action_lim = p->env == GCBF_ENV_DUBINS_CAR ? 2.f : 10.f;
  const double k = p->env == GCBF_ENV_DUBINS_CAR ? 3.0 : 4.0;
  c->safe_thr = (float)(k * R); c->warn_thr = (float)(k * R); c->coll_thr = (float)(2 * R);
  c->diag_safe = (float)(4 * R + 1); c->diag_unsafe = (float)((p->env == GCBF_ENV_SIMPLE_DRONE ? 2 : 4) * R + 1);
  c->two_r = (float)(R * 2);
  return GCBF_OK;
}

}  // namespace gcbf

using namespace gcbf;

extern "C" int gcbf_u_ref(const gcbf_env_cfg* cfg, const float* states, int ld_state, const float* goal, int ld_goal,
                          const float* K, float* u_ref, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_u_ref")) return rc;
  const int na = c.num_graphs * c.n;
  if (na == 0) return GCBF_OK;
  GCBF_REQUIRE(states && goal && u_ref && (K || c.env == GCBF_ENV_DUBINS_CAR), "gcbf_u_ref: null pointer");
  u_ref_kernel<<<ceil_div(na, 128), 128, 0, as_stream(stream)>>>(c, states, ld_state, goal, ld_goal, 0, K, u_ref);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

// the same with ONE GOAL SET PER GRAPH (goal [num_graphs * num_agents, ld_goal]): vectorised rollouts step many independent
// environments as one batch, each with its own goals
extern "C" int gcbf_u_ref_multi(const gcbf_env_cfg* cfg, const float* states, int ld_state, const float* goal, int ld_goal,
                                const float* K, float* u_ref, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_u_ref_multi")) return rc;
  const int na = c.num_graphs * c.n;
  if (na == 0) return GCBF_OK;
  GCBF_REQUIRE(states && goal && u_ref && (K || c.env == GCBF_ENV_DUBINS_CAR), "gcbf_u_ref_multi: null pointer");
  u_ref_kernel<<<ceil_div(na, 128), 128, 0, as_stream(stream)>>>(c, states, ld_state, goal, ld_goal, c.n, K, u_ref);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_step_fwd(const gcbf_env_cfg* cfg, const float* states, int ld_state, const float* action,
                             const float* goal, int ld_goal, const float* K, int freeze, float* states_next,
                             uint8_t* pass_mask, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_step_fwd")) return rc;
  const int64_t nn = (int64_t)c.num_graphs * c.N;
  if (nn == 0) return GCBF_OK;
  GCBF_REQUIRE(states && action && goal && states_next && pass_mask && (K || c.env == GCBF_ENV_DUBINS_CAR),
               "gcbf_step_fwd: null pointer");
  step_fwd_kernel<<<ceil_div(nn, 128), 128, 0, as_stream(stream)>>>(c, states, ld_state, action, goal, ld_goal, 0, K,
                                                                   freeze, states_next, pass_mask);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_step_fwd_multi(const gcbf_env_cfg* cfg, const float* states, int ld_state, const float* action,
                                   const float* goal, int ld_goal, const float* K, int freeze, float* states_next,
                                   uint8_t* pass_mask, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_step_fwd_multi")) return rc;
  const int64_t nn = (int64_t)c.num_graphs * c.N;
  if (nn == 0) return GCBF_OK;
  GCBF_REQUIRE(states && action && goal && states_next && pass_mask && (K || c.env == GCBF_ENV_DUBINS_CAR),
               "gcbf_step_fwd_multi: null pointer");
  step_fwd_kernel<<<ceil_div(nn, 128), 128, 0, as_stream(stream)>>>(c, states, ld_state, action, goal, ld_goal, c.n, K,
                                                                   freeze, states_next, pass_mask);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_step_bwd(const gcbf_env_cfg* cfg, const float* d_states_next, int ld_state,
                             const uint8_t* pass_mask, float* d_action, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_step_bwd")) return rc;
  const int na = c.num_graphs * c.n;
  if (na == 0) return GCBF_OK;
  GCBF_REQUIRE(d_states_next && pass_mask && d_action, "gcbf_step_bwd: null pointer");
  step_bwd_kernel<<<ceil_div(na, 128), 128, 0, as_stream(stream)>>>(c, d_states_next, ld_state, pass_mask, d_action);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

extern "C" int gcbf_masks(const gcbf_env_cfg* cfg, const float* states, int ld_state, uint8_t* safe, uint8_t* unsafe,
                          uint8_t* collision, void* stream) {
  EnvCfg c;
  if (int rc = make_cfg(cfg, &c, "gcbf_masks")) return rc;
  const int64_t na = (int64_t)c.num_graphs * c.n;
  if (na == 0) return GCBF_OK;
  GCBF_REQUIRE(states && safe && unsafe, "gcbf_masks: null pointer");
  masks_kernel<<<ceil_div(na * 32, 256), 256, 0, as_stream(stream)>>>(c, states, ld_state, safe, unsafe, collision);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}
