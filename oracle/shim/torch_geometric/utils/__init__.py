"""Shim of torch_geometric.utils.{softmax,index_to_mask,mask_to_index,to_networkx} (PyG 2.3.0)."""
import torch


def softmax(src, index=None, ptr=None, num_nodes=None, dim=0):
    """PyG 2.3 utils/softmax.py: per-group max of the DETACHED src is subtracted, exp, divided by
    (group sum + 1e-16)."""
    assert index is not None
    N = int(index.max()) + 1 if num_nodes is None else num_nodes
    dim = dim + src.dim() if dim < 0 else dim
    shape = list(src.shape)
    shape[dim] = N
    idx = index.view([-1 if i == dim else 1 for i in range(src.dim())]).expand_as(src)
    src_max = torch.full(shape, float('-inf'), dtype=src.dtype, device=src.device)
    src_max = src_max.scatter_reduce(dim, idx, src.detach(), reduce='amax', include_self=True)
    out = (src - src_max.gather(dim, idx)).exp()
    out_sum = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, idx, out) + 1e-16
    return out / out_sum.gather(dim, idx)


def index_to_mask(index, size=None):
    size = int(index.max()) + 1 if size is None else size
    mask = index.new_zeros(size, dtype=torch.bool)
    mask[index] = True
    return mask


def mask_to_index(mask):
    return mask.nonzero(as_tuple=False).view(-1)


def to_networkx(data, *a, **k):
    import networkx as nx
    g = nx.DiGraph()
    g.add_nodes_from(range(data.num_nodes))
    if data.edge_index is not None:
        g.add_edges_from(data.edge_index.t().tolist())
    return g
