"""Shim of torch_geometric.nn.conv.MessagePassing (PyG 2.3.0, flow='source_to_target', node_dim=-2):
x_j = x[edge_index[0]] (source), x_i = x[edge_index[1]] (target); aggregation index = edge_index[1],
dim_size = number of nodes.  Provides the private hooks gcbf/nn/gnn.py:44-53 and :101-104 reach into."""
import inspect
import torch


class _Inspector:
    def __init__(self, owner):
        self.owner = owner

    def distribute(self, func_name, kwargs):
        fn = getattr(self.owner, func_name)
        out = {}
        for name, p in inspect.signature(fn).parameters.items():
            if name in kwargs:
                out[name] = kwargs[name]
            elif p.default is not inspect.Parameter.empty:
                out[name] = p.default
        return out


class _MaxAggr(torch.nn.Module):
    def forward(self, x, index, ptr=None, dim_size=None, dim=-2):
        out = x.new_zeros((dim_size,) + tuple(x.shape[1:]))
        idx = index.view(-1, *([1] * (x.dim() - 1))).expand_as(x)
        return out.scatter_reduce(0, idx, x, reduce='amax', include_self=False)


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2, **kw):
        super().__init__()
        self.node_dim = node_dim
        if isinstance(aggr, torch.nn.Module):
            self.aggr_module = aggr
        elif aggr == 'max':
            self.aggr_module = _MaxAggr()
        elif aggr is None:
            self.aggr_module = None
        else:
            raise NotImplementedError(aggr)
        self.inspector = _Inspector(self)
        params = set(inspect.signature(self.message).parameters) | set(
            inspect.signature(self.update).parameters)
        self._user_args = sorted(params - {'aggr_out', 'inputs', 'self'})

    def _check_input(self, edge_index, size):
        return [None, None]

    def _collect(self, args, edge_index, size, kwargs):
        out = {}
        for arg in args:
            if arg.endswith('_i') or arg.endswith('_j'):
                data = kwargs.get(arg[:-2])
                if data is None:
                    continue
                if size[0] is None:
                    size[0] = size[1] = data.shape[self.node_dim]
                out[arg] = data.index_select(self.node_dim, edge_index[1 if arg.endswith('_i') else 0])
            elif arg in kwargs:
                out[arg] = kwargs[arg]
        out['edge_index'] = edge_index
        out['index'] = edge_index[1]
        out['ptr'] = None
        out['size'] = size
        out['dim_size'] = size[1]
        return out

    def propagate(self, edge_index, size=None, **kwargs):
        size = self._check_input(edge_index, size)
        coll = self._collect(self._user_args, edge_index, size, kwargs)
        msg = self.message(**self.inspector.distribute('message', coll))
        aggr_out = self.aggregate(msg, coll['index'], coll['ptr'], coll['dim_size'])
        upd_kwargs = self.inspector.distribute('update', coll)
        upd_kwargs.pop('aggr_out', None)
        return self.update(aggr_out, **upd_kwargs)

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        return self.aggr_module(inputs, index, ptr=ptr, dim_size=dim_size, dim=self.node_dim)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
