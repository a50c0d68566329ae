"""Host + device time to build one training batch from the replay storage: reference-style list of Data objects collated
by Batch.from_data_list (gcbf/algo/gcbf.py:149-159) vs the device-resident ring (gcbf_b200/algo/device_buffer.py)."""
import os, random, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gcbf-pytorch_b200'))
from gcbf_b200 import synth
from gcbf_b200.algo.device_buffer import collate
from gcbf_b200.data import Batch
dev = torch.device('cuda:0')
n, stored = 256, 240
env, algo = synth.seeded_algo('SimpleCar', n, dev, 0, {'num_obs': 0, 'area_size': 16.0})
_, ring_algo = synth.seeded_algo('SimpleCar', n, dev, 0, {'num_obs': 0, 'area_size': 16.0})
ring_algo.use_device_replay()
env.set_goal(synth.make_states('SimpleCar', n, 0, 1, 16.0, 999).goals)
for k in range(stored):
    sb = synth.make_states('SimpleCar', n, 0, 1, 16.0, 1000 + k)
    g = env.graph_from_states(sb.states.to(dev))
    algo.buffer.append(g, k % 4 != 0)
    ring_algo.buffer.append(g, k % 4 != 0)


def timed(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps):
        np.random.seed(i), random.seed(i)
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


t_list = timed(lambda: Batch.from_data_list(algo.buffer.sample(11, 3)))
t_ring = timed(lambda: collate(env, [(ring_algo.buffer, ring_algo.buffer.sample(11, 3))]))
np.random.seed(0), random.seed(0)
B = len(ring_algo.buffer.sample(11, 3))
print(f'batch of ~{B} graphs x {n} agents: list + Batch.from_data_list {t_list:.2f} ms, device ring + batched re-link {t_ring:.2f} ms')
