"""Vectorised rollout at C3 scale: ms per vector step and the edge count of every step.  `python tools/rollout_probe.py`"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ['bench.py']
import bench  # noqa: E402
from gcbf_b200.algo.rollout import VectorRollout  # noqa: E402

dev = torch.device('cuda', 0)
steps = int(os.environ.get('ROLL_STEPS', '10'))
print(json.dumps(bench.rollout_leg('C3', dev, steps=steps)))
sb, env, algo = bench.build_case('C3', dev, 0)
algo.use_device_replay(capacity=(steps + 4) * sb.num_graphs)
vr = VectorRollout(env, algo, sb.num_graphs, states=sb.states, goals=sb.goals.repeat(sb.num_graphs, 1))
for i in range(steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = vr.step(prob=0.5)
    torch.cuda.synchronize()
    print('step', i, 'edges', r['edge_count'], 'ms %.2f' % ((time.perf_counter() - t0) * 1e3))
