"""Shim of torch_geometric.data.{Data,Batch} (PyG 2.3.0 semantics used by the reference):
attribute store, None values dropped, `update`, `in`, num_nodes, Batch.from_data_list/to_data_list with
cumulative node offsets for keys containing "index" (PyG `__inc__`/`__cat_dim__` rules)."""
import copy
import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, '_store', {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos, **kwargs).items():
            setattr(self, k, v)

    # --- attribute protocol (PyG BaseStorage: None deletes / is not stored) -------------------------
    def __setattr__(self, key, value):
        if key.startswith('_'):
            object.__setattr__(self, key, value)
        elif value is None:
            self._store.pop(key, None)
        else:
            self._store[key] = value

    def __getattr__(self, key):
        store = object.__getattribute__(self, '_store')
        if key in store:
            return store[key]
        if key in ('x', 'edge_index', 'edge_attr', 'y', 'pos'):   # PyG properties: None when absent
            return None
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")

    def __delattr__(self, key):
        self._store.pop(key, None)

    def __contains__(self, key):
        return key in self._store

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def keys(self):
        return list(self._store.keys())

    def update(self, other):
        for k, v in other._store.items():
            setattr(self, k, v)
        return self

    @property
    def num_nodes(self):
        for k in ('x', 'pos', 'states'):
            if k in self._store:
                return self._store[k].shape[0]
        if 'edge_index' in self._store and self._store['edge_index'].numel():
            return int(self._store['edge_index'].max()) + 1
        return 0

    @property
    def num_edges(self):
        return self._store['edge_index'].shape[1] if 'edge_index' in self._store else 0

    def to(self, device):
        out = copy.copy(self)
        object.__setattr__(out, '_store', {k: (v.to(device) if torch.is_tensor(v) else v)
                                           for k, v in self._store.items()})
        return out

    def clone(self):
        out = copy.copy(self)
        object.__setattr__(out, '_store', {k: (v.clone() if torch.is_tensor(v) else copy.deepcopy(v))
                                           for k, v in self._store.items()})
        return out

    def __copy__(self):
        out = self.__class__.__new__(self.__class__)
        object.__setattr__(out, '_store', dict(self._store))
        for k, v in self.__dict__.items():
            if k != '_store':
                object.__setattr__(out, k, v)
        return out

    def __repr__(self):
        items = ', '.join(f'{k}={list(v.shape) if torch.is_tensor(v) else v}' for k, v in self._store.items())
        return f'{type(self).__name__}({items})'


class Batch(Data):
    """Block-diagonal collation (PyG 2.3 `collate`): tensors cat on dim 0, except keys containing
    'index' (cat on dim -1, incremented by the cumulative node count)."""

    @classmethod
    def from_data_list(cls, data_list):
        out = cls.__new__(cls)
        object.__setattr__(out, '_store', {})
        keys = list(data_list[0]._store.keys())
        num_nodes = [d.num_nodes for d in data_list]
        ptr = torch.zeros(len(data_list) + 1, dtype=torch.long)
        ptr[1:] = torch.tensor(num_nodes).cumsum(0)
        slices = {}
        for k in keys:
            vals = [d._store[k] for d in data_list]
            if torch.is_tensor(vals[0]):
                if 'index' in k:
                    dev = vals[0].device
                    vals2 = [v + int(ptr[i]) for i, v in enumerate(vals)]
                    out._store[k] = torch.cat(vals2, dim=-1)
                    sizes = [v.shape[-1] for v in vals]
                else:
                    out._store[k] = torch.cat(vals, dim=0)
                    sizes = [v.shape[0] for v in vals]
                s = torch.zeros(len(vals) + 1, dtype=torch.long)
                s[1:] = torch.tensor(sizes).cumsum(0)
                slices[k] = s
            else:
                out._store[k] = vals
                slices[k] = None
        dev = data_list[0]._store[keys[0]].device if keys else 'cpu'
        out._store['batch'] = torch.repeat_interleave(
            torch.arange(len(data_list)), torch.tensor(num_nodes)).to(dev)
        out._store['ptr'] = ptr.to(dev)
        object.__setattr__(out, '_slices', slices)
        object.__setattr__(out, '_num_graphs', len(data_list))
        object.__setattr__(out, '_node_ptr', ptr)
        return out

    @property
    def num_graphs(self):
        return self._num_graphs

    @property
    def num_nodes(self):
        return int(self._node_ptr[-1])

    def get_example(self, i):
        d = Data()
        for k, s in self._slices.items():
            v = self._store[k]
            if s is None:
                d._store[k] = v[i]
            elif 'index' in k:
                d._store[k] = v[..., int(s[i]):int(s[i + 1])] - int(self._node_ptr[i])
            else:
                d._store[k] = v[int(s[i]):int(s[i + 1])]
        return d

    def to_data_list(self):
        return [self.get_example(i) for i in range(self._num_graphs)]
