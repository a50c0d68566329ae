"""One GCBF.apply call (test-time controller) on one graph of a config, for `ncu` launch lists.  `python tools/one_apply.py C1 [max_iter]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C1'
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sys.argv = ['bench.py']
import bench  # noqa: E402

dev = torch.device('cuda', 0)
sb, env, algo = bench.build_case(cfg, dev, 0)
single = env.graph_from_states(sb.states[:sb.nodes_per_graph].to(dev))
a = algo.apply(single, max_iter=max_iter)
torch.cuda.synchronize()
print(cfg, 'rounds', algo.last_apply_rounds, 'max |action|', float(a.abs().max()))
