// TEST INFRASTRUCTURE ONLY: the kernels of csrc/macbf_kernels.cuh compiled as plain C++ and executed on an emulated 1-D grid
// (cuda_emu.h) -- the kernel bodies themselves (indexing, pitches, grid-stride loops), next to the per-element driver macbf_host.cpp.
#include "cuda_emu.h"
#include "macbf_kernels.cuh"

extern "C" {

void grid_radius_topk_count(int grid, int block, const float* states, int ld, int pos_dim, int num_graphs, int N, int n, float r, int metric, int k,
                            int32_t* counts) {
  EMU_LAUNCH(grid, block, gcbf::radius_topk_kernel<false>, states, ld, pos_dim, num_graphs, N, n, r, metric, k, counts, nullptr, nullptr, 0);
}

void grid_radius_topk_fill(int grid, int block, const float* states, int ld, int pos_dim, int num_graphs, int N, int n, float r, int metric, int k,
                           const int32_t* rowptr, int64_t* edge_index, int64_t E) {
  EMU_LAUNCH(grid, block, gcbf::radius_topk_kernel<true>, states, ld, pos_dim, num_graphs, N, n, r, metric, k, nullptr, rowptr, edge_index, E);
}

void grid_edge_masks(int grid, int block, const float* edge_attr, int ld, int pos_dim, int64_t E, float safe_thr, float coll_thr, uint8_t* safe,
                     uint8_t* unsafe) {
  EMU_LAUNCH(grid, block, gcbf::edge_masks_kernel, edge_attr, ld, pos_dim, E, safe_thr, coll_thr, safe, unsafe);
}

void grid_seg_max_fwd(int grid, int block, const float* msg, int ld_msg, const int32_t* rowptr, int num_nodes, int C, float* out, int ld_out,
                      int32_t* argmax) {
  EMU_LAUNCH(grid, block, gcbf::seg_max_fwd_kernel, msg, ld_msg, rowptr, num_nodes, C, out, ld_out, argmax);
}

void grid_seg_max_bwd(int grid, int block, const float* d_out, int ld_dout, const int32_t* argmax, int num_nodes, int C, float* d_msg, int ld_dmsg) {
  EMU_LAUNCH(grid, block, gcbf::seg_max_bwd_kernel, d_out, ld_dout, argmax, num_nodes, C, d_msg, ld_dmsg);
}

void grid_macbf_loss_grads(int grid, int block, const float* h, const float* hn, const uint8_t* safe, const uint8_t* unsafe, int64_t E, const float* act,
                           int ad, int64_t M, float alpha, float eps, float dt, float cu, float cs, float ch, float ca, const double* partial, float* d_h,
                           float* d_hn, float* d_act, float* scalars) {
  EMU_LAUNCH(grid, block, gcbf::macbf_loss_grads_kernel, h, hn, safe, unsafe, E, act, ad, M, alpha, eps, dt, cu, cs, ch, ca, partial, d_h, d_hn, d_act,
             scalars);
}

}  // extern "C"
