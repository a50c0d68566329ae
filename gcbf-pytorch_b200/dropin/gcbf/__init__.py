"""Drop-in alias: put `gcbf-pytorch_b200/dropin` (and `gcbf-pytorch_b200`) on sys.path and the reference's
import statements -- `from gcbf.nn import MLP`, `from gcbf.algo import make_algo`, `from gcbf.env import make_env`,
`from gcbf.controller import GNNController`, `from gcbf.trainer import Trainer` -- resolve to gcbf_b200."""
import importlib
import os
import sys

_pkg_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _pkg_root not in sys.path:
    sys.path.insert(0, _pkg_root)

import gcbf_b200 as _impl  # noqa: E402

for _name in ('nn', 'controller', 'algo', 'env', 'trainer', 'data'):
    _mod = importlib.import_module(f'gcbf_b200.{_name}')
    sys.modules[f'gcbf.{_name}'] = _mod
    globals()[_name] = _mod
for _sub in ('nn.mlp', 'nn.gnn', 'nn.utils', 'controller.gnn_controller', 'controller.macbf_controller', 'controller.nominal', 'controller.base', 'algo.gcbf', 'algo.macbf',
             'algo.nominal', 'algo.base',
             'algo.buffer', 'env.base', 'env.simple_car', 'env.dubins_car', 'env.simple_drone', 'trainer.trainer',
             'trainer.utils'):
    sys.modules[f'gcbf.{_sub}'] = importlib.import_module(f'gcbf_b200.{_sub}')
