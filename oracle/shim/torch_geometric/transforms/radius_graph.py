"""Shim of torch_geometric.transforms.RadiusGraph -> torch_cluster.radius_graph (restated from the
published algorithm, torch_cluster csrc/cuda/radius_cuda.cu / csrc/cpu/radius_cpu.cpp):
for every query i keep points j with SQUARED distance sum_d (x_i[d]-x_j[d])^2 < r*r (strict; sequential,
unfused accumulate over d in fp32), at most max_num_neighbors+1 hits including self in ascending j
(the CUDA kernel's order; the CPU nanoflann order is unsorted, so (i asc, j asc) is the canonical order),
then drop self loops.  edge_index = [source j ; target i] grouped by target ascending."""
import numpy as np
import torch


class RadiusGraph:
    def __init__(self, r, loop=False, max_num_neighbors=32, flow='source_to_target', num_workers=1):
        self.r, self.loop, self.max_num_neighbors = r, loop, max_num_neighbors

    def __call__(self, data):
        data.edge_attr = None
        pos = data.pos
        P = pos.detach().cpu().numpy().astype(np.float32)
        n = P.shape[0]
        d2 = np.zeros((n, n), dtype=np.float32)
        for d in range(P.shape[1]):                       # sequential fp32 accumulate, no FMA
            diff = (P[:, None, d] - P[None, :, d]).astype(np.float32)
            d2 = (d2 + (diff * diff).astype(np.float32)).astype(np.float32)
        hit = d2 < np.float32(np.float32(self.r) * np.float32(self.r))
        # cap at max_num_neighbors + 1 hits (self included) in ascending j
        order = np.cumsum(hit, axis=1)
        hit &= order <= (self.max_num_neighbors + (0 if self.loop else 1))
        if not self.loop:
            hit[np.arange(n), np.arange(n)] = False
        i, j = np.nonzero(hit)                            # row-major: i asc, j asc
        ei = torch.from_numpy(np.stack([j, i]).astype(np.int64)).to(pos.device)
        data.edge_index = ei
        return data
