"""Minimal graph containers with the slice of torch_geometric.data.{Data, Batch} behaviour the reference
relies on (torch_geometric itself is not a dependency here):

  * keyword construction, attribute access, `hasattr` false for absent keys, `None` removes a key,
    `update(other)`, `key in data`, `num_nodes`;
  * `Batch.from_data_list` = block-diagonal collation (reference call sites gcbf/algo/gcbf.py:159,200):
    tensors are concatenated along dim 0, except keys containing "index" which are concatenated along the
    last dim after adding the running node offset; `to_data_list()` undoes it.

Anything that quacks the same way (e.g. a real torch_geometric Data) is accepted by the modules; these
classes exist so the package works without PyG.
"""
import weakref
from typing import Dict, List, Optional

import torch

_OPTIONAL_NONE = ('x', 'edge_index', 'edge_attr', 'y', 'pos')   # PyG exposes these as None when absent


class Data:
    def __init__(self, **fields):
        object.__setattr__(self, '_fields', {})
        for name, value in fields.items():
            self[name] = value

    # -- mapping protocol ---------------------------------------------------------------------------
    def __setitem__(self, name: str, value):
        if value is None:
            self._fields.pop(name, None)
        else:
            self._fields[name] = value

    def __getitem__(self, name: str):
        return self._fields[name]

    def __contains__(self, name: str) -> bool:
        return name in self._fields

    def keys(self) -> List[str]:
        return list(self._fields)

    # -- attribute protocol -------------------------------------------------------------------------
    def __getattr__(self, name: str):
        fields = object.__getattribute__(self, '_fields')
        if name in fields:
            return fields[name]
        if name in _OPTIONAL_NONE:
            return None
        raise AttributeError(f'{type(self).__name__} has no field {name!r}')

    def __setattr__(self, name: str, value):
        if name.startswith('_'):
            object.__setattr__(self, name, value)
        else:
            self[name] = value

    def __delattr__(self, name: str):
        self._fields.pop(name, None)

    def update(self, other: 'Data') -> 'Data':
        for name in other.keys():
            self[name] = other[name]
        return self

    @property
    def num_nodes(self) -> int:
        for name in ('x', 'pos', 'states'):
            if name in self._fields:
                return int(self._fields[name].shape[0])
        return 0

    @property
    def num_edges(self) -> int:
        return int(self._fields['edge_index'].shape[1]) if 'edge_index' in self._fields else 0

    def to(self, device) -> 'Data':
        out = type(self).__new__(type(self))
        object.__setattr__(out, '_fields', {k: (v.to(device) if torch.is_tensor(v) else v)
                                            for k, v in self._fields.items()})
        for k, v in self.__dict__.items():
            if k != '_fields':
                object.__setattr__(out, k, v)
        return out

    def __repr__(self) -> str:
        body = ', '.join(f'{k}={tuple(v.shape) if torch.is_tensor(v) else v!r}' for k, v in self._fields.items())
        return f'{type(self).__name__}({body})'


class Batch(Data):
    """Collated graphs.  Keeps the per-key split points so `to_data_list()` can restore the inputs."""

    @classmethod
    def from_data_list(cls, graphs: List[Data]) -> 'Batch':
        if len(graphs) == 0:
            raise ValueError('from_data_list needs at least one graph')
        out = cls()
        sizes = [g.num_nodes for g in graphs]
        offsets = [0]
        for s in sizes:
            offsets.append(offsets[-1] + s)
        splits: Dict[str, Optional[List[int]]] = {}
        for name in graphs[0].keys():
            parts = [g[name] for g in graphs]
            if not torch.is_tensor(parts[0]):
                out[name] = parts
                splits[name] = None
                continue
            if 'index' in name:
                out[name] = torch.cat([p + off for p, off in zip(parts, offsets)], dim=-1)
                lens = [int(p.shape[-1]) for p in parts]
            else:
                out[name] = torch.cat(parts, dim=0)
                lens = [int(p.shape[0]) for p in parts]
            acc = [0]
            for n in lens:
                acc.append(acc[-1] + n)
            splits[name] = acc
        device = next((v.device for v in out._fields.values() if torch.is_tensor(v)), torch.device('cpu'))
        out['batch'] = torch.repeat_interleave(torch.arange(len(graphs), device=device),
                                               torch.tensor(sizes, device=device))
        out['ptr'] = torch.tensor(offsets, device=device)
        object.__setattr__(out, '_splits', splits)
        object.__setattr__(out, '_offsets', offsets)
        return out

    @property
    def num_graphs(self) -> int:
        return len(self._offsets) - 1

    @property
    def num_nodes(self) -> int:
        return self._offsets[-1]

    def get_example(self, i: int) -> Data:
        g = Data()
        for name, acc in self._splits.items():
            value = self._fields[name]
            if acc is None:
                g[name] = value[i]
            elif 'index' in name:
                g[name] = value[..., acc[i]:acc[i + 1]] - self._offsets[i]
            else:
                g[name] = value[acc[i]:acc[i + 1]]
        return g

    def to_data_list(self) -> List[Data]:
        return [self.get_example(i) for i in range(self.num_graphs)]


_INDEX_CACHE = {}


def agent_row_index(data) -> Optional[torch.Tensor]:
    """int64 row indices selected by `data.agent_mask` (None when the graph has no mask, i.e. SimpleCar where
    every node is an agent) -- the `x[data.agent_mask]` of reference gcbf/algo/gcbf.py:52-53.  Cached per mask
    tensor (nonzero() synchronises)."""
    if not hasattr(data, 'agent_mask'):
        return None
    mask = data.agent_mask
    if mask is None:
        return None
    key = (id(mask), mask.data_ptr(), mask.numel(), mask._version)
    hit = _INDEX_CACHE.get(key)
    if hit is not None and hit[0]() is mask:
        return hit[1]
    idx = torch.nonzero(mask, as_tuple=False).reshape(-1)
    if len(_INDEX_CACHE) > 64:
        _INDEX_CACHE.clear()
    _INDEX_CACHE[key] = (weakref.ref(mask), idx)
    return idx
