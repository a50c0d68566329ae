// Kernels of the MACBF path that are plain grid-stride loops around the per-element functions of macbf_core.h: top-k
// filtered radius graph, per-edge masks, max aggregation forward / backward, loss gradients.  In a header of their own, free of CUDA
// runtime includes, so that tests/host_driver/macbf_grid.cpp can compile the SAME kernel bodies for the host on an emulated grid
// (tests/host_driver/cuda_emu.h).  (The partial-sum kernel, which needs shared memory / shuffles / atomics, stays in macbf.cu.)
#pragma once
#include "gcbf_b200.h"
#include "macbf_core.h"

namespace gcbf {

using namespace macbf;

// one thread per target agent: count (FILL = false) or write (FILL = true) its kept neighbours in ascending source order
template <bool FILL>
__global__ void radius_topk_kernel(const float* __restrict__ states, int ld, int pos_dim, int num_graphs, int N, int n, float r,
                                   int metric, int k, int32_t* __restrict__ counts, const int32_t* __restrict__ rowptr,
                                   int64_t* __restrict__ edge_index, int64_t E) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)num_graphs * n) return;
  const int g = (int)(t / n), il = (int)(t % n);
  const int64_t node_base = (int64_t)g * N;
  const float* base = states + node_base * ld;
  if (FILL) {
    const int64_t off = rowptr[t];
    topk_row(base, ld, pos_dim, N, il, r, k, metric, node_base, edge_index + off, edge_index + E + off);
  } else {
    counts[t] = topk_row(base, ld, pos_dim, N, il, r, k, metric, node_base, nullptr, nullptr);
  }
}

__global__ void edge_masks_kernel(const float* __restrict__ edge_attr, int ld, int pos_dim, int64_t E, float safe_thr, float coll_thr,
                                  uint8_t* __restrict__ safe, uint8_t* __restrict__ unsafe) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    float row[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < pos_dim; ++k) row[k] = edge_attr[e * ld + k];
    uint8_t s, u;
    edge_flags(row, pos_dim, safe_thr, coll_thr, &s, &u);
    safe[e] = s;
    unsafe[e] = u;
  }
}

// thread per (node, channel), channel fastest: the loads of one edge row are coalesced across the threads of a node
__global__ void seg_max_fwd_kernel(const float* __restrict__ msg, int ld_msg, const int32_t* __restrict__ rowptr, int num_nodes, int C,
                                   float* __restrict__ out, int ld_out, int32_t* __restrict__ argmax) {
  const int64_t total = (int64_t)num_nodes * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / C), c = (int)(idx % C);
    float v;
    int32_t a;
    seg_max_cell(msg, ld_msg, rowptr[i], rowptr[i + 1], c, &v, &a);
    out[(int64_t)i * ld_out + c] = v;
    argmax[idx] = a;
  }
}

// d_msg was zeroed; every (edge, channel) cell is the argmax of at most one (node, channel) cell, so plain stores suffice
__global__ void seg_max_bwd_kernel(const float* __restrict__ d_out, int ld_dout, const int32_t* __restrict__ argmax, int num_nodes, int C,
                                   float* __restrict__ d_msg, int ld_dmsg) {
  const int64_t total = (int64_t)num_nodes * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int32_t a = argmax[idx];
    if (a < 0) continue;
    const int i = (int)(idx / C), c = (int)(idx % C);
    d_msg[(int64_t)a * ld_dmsg + c] = d_out[(int64_t)i * ld_dout + c];
  }
}

__global__ void macbf_loss_grads_kernel(const float* __restrict__ h, const float* __restrict__ hn, const uint8_t* __restrict__ safe,
                                        const uint8_t* __restrict__ unsafe, int64_t E, const float* __restrict__ act, int ad, int64_t M,
                                        float alpha, float eps, float dt, float cu, float cs, float ch, float ca,
                                        const double* __restrict__ partial, float* __restrict__ d_h, float* __restrict__ d_hn,
                                        float* __restrict__ d_act, float* __restrict__ scalars) {
  const double cnt_u = partial[MLP_CNT_UNSAFE], cnt_s = partial[MLP_CNT_SAFE], cnt_e = partial[MLP_CNT_EDGES], cnt_a = partial[MLP_CNT_AGENTS];
  const float inv_u = cnt_u > 0 ? (float)(1.0 / cnt_u) : 0.f;
  const float inv_s = cnt_s > 0 ? (float)(1.0 / cnt_s) : 0.f;
  const float inv_e = cnt_e > 0 ? (float)(1.0 / cnt_e) : 0.f;
  const float inv_a = cnt_a > 0 ? (float)(1.0 / cnt_a) : 0.f;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {
    const float lu = cnt_u > 0 ? (float)(partial[MLP_SUM_UNSAFE] / cnt_u) : 0.f;     // empty mask: loss 0, accuracy 1 (macbf.py:152-153)
    const float ls = cnt_s > 0 ? (float)(partial[MLP_SUM_SAFE] / cnt_s) : 0.f;
    const float lh = cnt_e > 0 ? (float)(partial[MLP_SUM_HDOT] / cnt_e) : 0.f;
    const float la = cnt_a > 0 ? (float)(partial[MLP_SUM_ACT] / cnt_a) : 0.f;
    scalars[0] = lu; scalars[1] = ls; scalars[2] = lh; scalars[3] = la;
    scalars[4] = cnt_u > 0 ? (float)(partial[MLP_OK_UNSAFE] / cnt_u) : 1.f;
    scalars[5] = cnt_s > 0 ? (float)(partial[MLP_OK_SAFE] / cnt_s) : 1.f;
    scalars[6] = cu * lu + cs * ls + ch * lh + ca * la;                                // macbf.py:174-177
    scalars[7] = cnt_e > 0 ? (float)(partial[MLP_OK_HDOT] / cnt_e) : 1.f;             // acc/derivative (macbf.py:168)
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = t; e < E; e += stride) {
    float g, gn;
    edge_grads(h[e], hn[e], safe[e], unsafe[e], alpha, eps, dt, cu, cs, ch, inv_u, inv_s, inv_e, &g, &gn);
    d_h[e] = g;
    d_hn[e] = gn;
  }
  for (int64_t i = t; i < M * ad; i += stride) d_act[i] = ca * inv_a * 2.f * act[i];
}

}  // namespace gcbf
