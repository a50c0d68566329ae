"""Shim of torch_geometric.nn.Sequential (children are named module_{i}: state-dict key contract)."""
import torch
from . import conv, aggr  # noqa: F401


class Sequential(torch.nn.Module):
    def __init__(self, input_args, modules):
        super().__init__()
        self._in = [a.strip() for a in input_args.split(',')]
        self._specs = []
        for i, m in enumerate(modules):
            if isinstance(m, (tuple, list)):
                mod, desc = m
                ins, outs = desc.split('->')
                ins = [a.strip() for a in ins.split(',')]
                outs = [a.strip() for a in outs.split(',')]
            else:
                mod, ins, outs = m, None, None
            setattr(self, f'module_{i}', mod)
            self._specs.append((f'module_{i}', ins, outs))

    def forward(self, *args):
        env = dict(zip(self._in, args))
        last = None
        for name, ins, outs in self._specs:
            mod = getattr(self, name)
            if ins is None:
                last = mod(last)
            else:
                res = mod(*[env[a] for a in ins])
                if len(outs) == 1:
                    env[outs[0]] = res
                else:
                    for o, r in zip(outs, res):
                        env[o] = r
                last = res
        return last
