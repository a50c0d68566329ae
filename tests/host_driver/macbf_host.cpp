// TEST INFRASTRUCTURE ONLY: host build of the per-element functions the MACBF kernels are made of
// (gcbf-pytorch_b200/csrc/macbf_core.h), so that the CPU test-suite -- which has no GPU -- checks the arithmetic the device
// executes against the reference-on-shim.  Compiled by tests/test_macbf_cpu.py with `g++ -O2 -ffp-contract=off -shared -fPIC`.
// Each function is the serial form of the corresponding kernel's grid-stride loop in csrc/macbf.cu.
#include <stdint.h>
#include <string.h>
#include "macbf_core.h"

using namespace gcbf::macbf;

extern "C" {

// rowptr[0 .. num_graphs * n] (exclusive scan of the per-target counts), then the edges: returns E; edge_index may be null
int64_t host_radius_graph_topk(const float* states, int ld, int pos_dim, int num_graphs, int N, int n, float r, int metric, int k,
                               int32_t* rowptr, int64_t* edge_index, int64_t capacity) {
  const int64_t na = (int64_t)num_graphs * n;
  int64_t total = 0;
  for (int64_t t = 0; t < na; ++t) {
    const int g = (int)(t / n), il = (int)(t % n);
    rowptr[t] = (int32_t)total;
    total += topk_row(states + (int64_t)g * N * ld, ld, pos_dim, N, il, r, k, metric, (int64_t)g * N, nullptr, nullptr);
  }
  rowptr[na] = (int32_t)total;
  if (edge_index == nullptr || total > capacity) return total;
  for (int64_t t = 0; t < na; ++t) {
    const int g = (int)(t / n), il = (int)(t % n);
    topk_row(states + (int64_t)g * N * ld, ld, pos_dim, N, il, r, k, metric, (int64_t)g * N, edge_index + rowptr[t],
             edge_index + total + rowptr[t]);
  }
  return total;
}

void host_edge_masks(const float* edge_attr, int ld, int pos_dim, int64_t E, double agent_radius, uint8_t* safe, uint8_t* unsafe) {
  const float safe_thr = (float)(4 * agent_radius), coll_thr = (float)(2 * agent_radius);
  for (int64_t e = 0; e < E; ++e) edge_flags(edge_attr + e * ld, pos_dim, safe_thr, coll_thr, safe + e, unsafe + e);
}

void host_seg_max_fwd(const float* msg, int ld_msg, const int32_t* rowptr, int num_nodes, int C, float* out, int ld_out, int32_t* argmax) {
  for (int i = 0; i < num_nodes; ++i)
    for (int c = 0; c < C; ++c) seg_max_cell(msg, ld_msg, rowptr[i], rowptr[i + 1], c, out + (int64_t)i * ld_out + c, argmax + (int64_t)i * C + c);
}

void host_seg_max_bwd(const float* d_out, int ld_dout, const int32_t* argmax, int num_nodes, int C, float* d_msg, int ld_dmsg, int64_t E) {
  memset(d_msg, 0, sizeof(float) * (size_t)E * ld_dmsg);
  for (int i = 0; i < num_nodes; ++i)
    for (int c = 0; c < C; ++c) {
      const int32_t a = argmax[(int64_t)i * C + c];
      if (a >= 0) d_msg[(int64_t)a * ld_dmsg + c] = d_out[(int64_t)i * ld_dout + c];
    }
}

// the two passes of gcbf_macbf_loss_partials / gcbf_macbf_loss_grads (ranks may all-reduce `partial` in between)
void host_macbf_loss_partials(const float* h, const float* hn, const uint8_t* safe, const uint8_t* unsafe, int64_t E, const float* act, int ad,
                              int64_t M, float alpha, float eps, float dt, double* partial) {
  for (int k = 0; k < MLP_SIZE; ++k) partial[k] = 0.0;
  for (int64_t e = 0; e < E; ++e) edge_terms(h[e], hn[e], safe[e], unsafe[e], alpha, eps, dt, partial);
  for (int64_t i = 0; i < M; ++i) { partial[MLP_SUM_ACT] += action_term(act + i * ad, ad); partial[MLP_CNT_AGENTS] += 1.0; }
}

void host_macbf_loss_grads(const float* h, const float* hn, const uint8_t* safe, const uint8_t* unsafe, int64_t E, const float* act, int ad, int64_t M,
                           float alpha, float eps, float dt, float cu, float cs, float ch, float ca, const double* partial, float* d_h, float* d_hn,
                           float* d_act, float* scalars) {
  const double cnt_u = partial[MLP_CNT_UNSAFE], cnt_s = partial[MLP_CNT_SAFE], cnt_e = partial[MLP_CNT_EDGES], cnt_a = partial[MLP_CNT_AGENTS];
  const float inv_u = cnt_u > 0 ? (float)(1.0 / cnt_u) : 0.f, inv_s = cnt_s > 0 ? (float)(1.0 / cnt_s) : 0.f;
  const float inv_e = cnt_e > 0 ? (float)(1.0 / cnt_e) : 0.f, inv_a = cnt_a > 0 ? (float)(1.0 / cnt_a) : 0.f;
  const float lu = cnt_u > 0 ? (float)(partial[MLP_SUM_UNSAFE] / cnt_u) : 0.f, ls = cnt_s > 0 ? (float)(partial[MLP_SUM_SAFE] / cnt_s) : 0.f;
  const float lh = cnt_e > 0 ? (float)(partial[MLP_SUM_HDOT] / cnt_e) : 0.f, la = cnt_a > 0 ? (float)(partial[MLP_SUM_ACT] / cnt_a) : 0.f;
  scalars[0] = lu; scalars[1] = ls; scalars[2] = lh; scalars[3] = la;
  scalars[4] = cnt_u > 0 ? (float)(partial[MLP_OK_UNSAFE] / cnt_u) : 1.f;
  scalars[5] = cnt_s > 0 ? (float)(partial[MLP_OK_SAFE] / cnt_s) : 1.f;
  scalars[6] = cu * lu + cs * ls + ch * lh + ca * la;
  scalars[7] = cnt_e > 0 ? (float)(partial[MLP_OK_HDOT] / cnt_e) : 1.f;
  for (int64_t e = 0; e < E; ++e) edge_grads(h[e], hn[e], safe[e], unsafe[e], alpha, eps, dt, cu, cs, ch, inv_u, inv_s, inv_e, d_h + e, d_hn + e);
  for (int64_t i = 0; i < M * ad; ++i) d_act[i] = ca * inv_a * 2.f * act[i];
}

}  // extern "C"
