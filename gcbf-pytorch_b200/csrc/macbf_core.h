// Per-element arithmetic of the MACBF path, written once for device and host.
//
// The kernels in macbf.cu are grid-stride loops around these functions; tests/host_driver/macbf_host.cpp compiles the SAME
// functions with gcc (-ffp-contract=off) so their results are checked against the reference-on-shim in the CPU test suite,
// where no GPU exists.  On the device every fp32 operation is an explicit round-to-nearest intrinsic (no contraction unless the
// reference's CPU arithmetic contracts, see pair_dist); on the host the plain operators compile to the same IEEE operations.
//
// Reference sites:
//   top-k neighbour filter      gcbf/env/dubins_car.py:736-740, simple_drone.py:322-326 (torch.topk on the masked distance rows),
//                               simple_car.py:32-33 (RadiusGraph(max_num_neighbors=k) -> torch_cluster's first k+1 hits)
//   per-edge safe / unsafe      simple_car.py:307-311, 332-336; dubins_car.py:819-823, 844-848; simple_drone.py:380-384, 405-409
//   max aggregation             gcbf/nn/gnn.py:116-119 (MessagePassing(aggr='max'))
//   losses                      gcbf/algo/macbf.py:140-181
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GCBF_HD __host__ __device__ __forceinline__
#else
#define GCBF_HD inline
#endif

namespace gcbf {
namespace macbf {

GCBF_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
GCBF_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
GCBF_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
GCBF_HD float f_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}
GCBF_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
GCBF_HD float f_sqrt(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}

// torch.norm(dim=-1) of a 2- or 3-vector as torch's CPU kernel forms it: acc = fma(d, d, acc) over the dims, then sqrt
// (the same formula graph.cu / env.cu use for the radius graph and the node masks; measured there against torch 2.11)
GCBF_HD float norm_fma(const float* d, int n) {
  float acc = 0.f;
  for (int k = 0; k < n; ++k) acc = f_fma(d[k], d[k], acc);
  return f_sqrt(acc);
}

// ||pos_i - pos_j|| (metric 1: DubinsCar / SimpleDrone, dense torch.norm path)
GCBF_HD float pair_dist(const float* pi, const float* pj, int pos_dim) {
  float d[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < pos_dim; ++k) d[k] = f_sub(pi[k], pj[k]);
  return norm_fma(d, pos_dim);
}

// squared distance as torch_cluster accumulates it (metric 0: SimpleCar): sequential, unfused
GCBF_HD float pair_d2(const float* pi, const float* pj, int pos_dim) {
  float d2 = 0.f;
  for (int k = 0; k < pos_dim; ++k) {
    const float diff = f_sub(pi[k], pj[k]);
    d2 = f_add(d2, f_mul(diff, diff));
  }
  return d2;
}

// Neighbours of agent `il` (local index inside its graph) under the top-k filter.  `base` points at the first state row of the
// graph, rows are `ld` floats apart, the first `pos_dim` columns are the position, the graph has N nodes (agents first).
// Returns the number of kept sources; if src_out is not null the kept GLOBAL source ids (node_base + j) are written in ascending j
// together with the target id.
//
// metric 0 (SimpleCar): torch_cluster.radius keeps the first k+1 hits of `d2 < r*r` in ascending j, the query point itself
//   included (PyG calls it with max_num_neighbors + 1 when loop=False), and PyG then drops the self loop.
// metric 1 (DubinsCar / SimpleDrone): the k smallest entries of the distance row (diagonal + r + 1, so never selected ahead of an
//   in-radius node) keep their distance, every other entry is pushed beyond the radius; then `dist < r`.  Hence: if at most k
//   nodes are inside the radius all of them are kept, otherwise the k nearest.  torch.topk leaves the order of EQUAL distances
//   unspecified; here the lower index wins.
GCBF_HD int topk_row(const float* base, int ld, int pos_dim, int N, int il, float r, int k, int metric, int64_t node_base,
                     int64_t* src_out, int64_t* dst_out) {
  float pi[3] = {0.f, 0.f, 0.f};
  for (int d = 0; d < pos_dim; ++d) pi[d] = base[(int64_t)il * ld + d];
  int kept = 0;
  if (metric == 0) {
    const float r2 = f_mul(r, r);
    int hits = 0;
    for (int j = 0; j < N; ++j) {
      if (!(pair_d2(pi, base + (int64_t)j * ld, pos_dim) < r2)) continue;
      if (hits >= k + 1) break;
      ++hits;
      if (j == il) continue;
      if (src_out) { src_out[kept] = node_base + j; dst_out[kept] = node_base + il; }
      ++kept;
    }
    return kept;
  }
  int inside = 0;
  for (int j = 0; j < N; ++j)
    if (j != il && pair_dist(pi, base + (int64_t)j * ld, pos_dim) < r) ++inside;
  for (int j = 0; j < N; ++j) {
    if (j == il) continue;
    const float dj = pair_dist(pi, base + (int64_t)j * ld, pos_dim);
    if (!(dj < r)) continue;
    bool keep = true;
    if (inside > k) {
      int rank = 0;                                  // entries of the row ahead of j in (distance, index) order
      for (int j2 = 0; j2 < N && rank < k; ++j2) {
        if (j2 == il || j2 == j) continue;
        const float d2 = pair_dist(pi, base + (int64_t)j2 * ld, pos_dim);
        if (d2 < dj || (d2 == dj && j2 < j)) ++rank;
      }
      keep = rank < k;
    }
    if (!keep) continue;
    if (src_out) { src_out[kept] = node_base + j; dst_out[kept] = node_base + il; }
    ++kept;
  }
  return kept;
}

// per-edge masks: dist = ||edge_attr[:pos_dim]||; safe = dist > 4R, unsafe (collision) = dist < 2R.  All three envs use 4R / 2R here
// (the node-level masks differ per env; the edge-level ones do not).
GCBF_HD void edge_flags(const float* edge_attr_row, int pos_dim, float safe_thr, float coll_thr, uint8_t* safe, uint8_t* unsafe) {
  const float dist = norm_fma(edge_attr_row, pos_dim);
  *safe = dist > safe_thr ? 1 : 0;
  *unsafe = dist < coll_thr ? 1 : 0;
}

// max over the incoming messages of one (node, channel) cell; empty neighbourhoods give 0 (PyG fills missing groups with 0) and
// argmax -1.  First maximum wins on ties.
GCBF_HD void seg_max_cell(const float* msg, int ld_msg, int beg, int end, int c, float* val, int32_t* arg) {
  if (beg >= end) { *val = 0.f; *arg = -1; return; }
  float best = msg[(int64_t)beg * ld_msg + c];
  int32_t a = beg;
  for (int e = beg + 1; e < end; ++e) {
    const float v = msg[(int64_t)e * ld_msg + c];
    if (v > best) { best = v; a = e; }
  }
  *val = best;
  *arg = a;
}

// ---- losses (macbf.py:140-181) -----------------------------------------------------------------------------------------------
// partial sums, double[GCBF_MLP_SIZE]; indices 0..8 as the GCBF losses (include/gcbf_b200.h GCBF_LP_*), plus
enum { MLP_SUM_UNSAFE = 0, MLP_CNT_UNSAFE = 1, MLP_OK_UNSAFE = 2, MLP_SUM_SAFE = 3, MLP_CNT_SAFE = 4, MLP_OK_SAFE = 5,
       MLP_SUM_HDOT = 6, MLP_CNT_EDGES = 7, MLP_SUM_ACT = 8, MLP_OK_HDOT = 9, MLP_CNT_AGENTS = 10, MLP_SIZE = 16 };

GCBF_HD float hdot_of(float h, float hn, float dt) { return f_div(f_sub(hn, h), dt); }                       // macbf.py:165
GCBF_HD float hdot_arg(float h, float hd, float alpha, float eps) {                                          // macbf.py:166
  return f_add(f_sub(-hd, f_mul(alpha, h)), eps);
}

GCBF_HD void edge_terms(float h, float hn, uint8_t safe, uint8_t unsafe, float alpha, float eps, float dt, double* acc) {
  if (unsafe) {                                            // macbf.py:144-150
    acc[MLP_SUM_UNSAFE] += fmaxf(f_add(h, eps), 0.f);
    acc[MLP_CNT_UNSAFE] += 1.0;
    acc[MLP_OK_UNSAFE] += (h < 0.f) ? 1.0 : 0.0;
  }
  if (safe) {                                              // macbf.py:156-161
    acc[MLP_SUM_SAFE] += fmaxf(f_add(-h, eps), 0.f);
    acc[MLP_CNT_SAFE] += 1.0;
    acc[MLP_OK_SAFE] += (h >= 0.f) ? 1.0 : 0.0;
  }
  const float hd = hdot_of(h, hn, dt);
  acc[MLP_SUM_HDOT] += fmaxf(hdot_arg(h, hd, alpha, eps), 0.f);
  acc[MLP_OK_HDOT] += (f_add(hd, f_mul(alpha, h)) >= 0.f) ? 1.0 : 0.0;      // macbf.py:168
  acc[MLP_CNT_EDGES] += 1.0;
}

GCBF_HD float action_term(const float* act_row, int ad) {  // macbf.py:171
  float s = 0.f;
  for (int k = 0; k < ad; ++k) s = f_add(s, f_mul(act_row[k], act_row[k]));
  return s;
}

// d loss / d h and d loss / d h_next of one edge; inv_* = 1 / (global) count or 0 for an empty mask
GCBF_HD void edge_grads(float h, float hn, uint8_t safe, uint8_t unsafe, float alpha, float eps, float dt, float cu, float cs, float ch,
                        float inv_u, float inv_s, float inv_e, float* g_h, float* g_hn) {
  float g = 0.f, gn = 0.f;
  if (unsafe && f_add(h, eps) > 0.f) g += cu * inv_u;
  if (safe && f_add(-h, eps) > 0.f) g -= cs * inv_s;
  const float hd = hdot_of(h, hn, dt);
  if (hdot_arg(h, hd, alpha, eps) > 0.f) {
    const float w = ch * inv_e;                            // d relu(-(hn - h)/dt - alpha h + eps): d/dh = 1/dt - alpha, d/dhn = -1/dt
    g += w / dt - w * alpha;
    gn = -(w / dt);
  }
  *g_h = g;
  *g_hn = gn;
}

}  // namespace macbf
}  // namespace gcbf
