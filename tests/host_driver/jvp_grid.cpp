// TEST INFRASTRUCTURE ONLY: the kernels of csrc/jvp_kernels.cuh compiled as plain C++ and executed on an emulated 1-D grid
// (cuda_emu.h): checks the kernel bodies themselves -- indexing, pitches, the grid-stride loops -- which the per-element host driver
// (jvp_host.cpp) does not contain.  Compiled by tests/test_jvp_cpu.py with g++ -ffp-contract=off.
#include "cuda_emu.h"
#include "jvp_kernels.cuh"

extern "C" {

void grid_state_dot(int grid, int block, int env, int num_graphs, int N, int n, const float* states, int ld, const float* action, const float* u_ref,
                    const float* goal, int ld_goal, int goal_gstride, float action_lim, float speed_limit, float dist2goal, int freeze, float* out,
                    int ld_out) {
  EMU_LAUNCH(grid, block, gcbf::state_dot_kernel, env, num_graphs, N, n, states, ld, action, u_ref, goal, ld_goal, goal_gstride, action_lim,
             speed_limit, dist2goal, freeze, out, ld_out);
}

void grid_edge_attr_tangent(int grid, int block, int env, const float* states, int ld, const float* sdot, int ld_sd, const int64_t* ei, int64_t E,
                            float* out) {
  EMU_LAUNCH(grid, block, gcbf::edge_attr_tangent_kernel, env, states, ld, sdot, ld_sd, ei, E, out);
}

void grid_attn_tangent(int grid, int block, const float* msg, int ld_msg, const float* t_msg, int ld_tmsg, const float* att, const float* t_gate,
                       const int32_t* rowptr, int num_nodes, int C, float* out, int ld_out) {
  EMU_LAUNCH(grid, block, gcbf::attn_tangent_kernel, msg, ld_msg, t_msg, ld_tmsg, att, t_gate, rowptr, num_nodes, C, out, ld_out);
}

}  // extern "C"
