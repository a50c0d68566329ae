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

# where does a slow step spend its time?  wrap the phases of VectorRollout.step with synchronising timers
import collections  # noqa: E402
phase = collections.OrderedDict()


def timed(name, fn):
    def w(*a, **k):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        phase[name] = phase.get(name, 0.0) + (time.perf_counter() - t) * 1e3
        return r
    return w


env.add_communication_links = timed('links', env.add_communication_links)
env.make_graph = timed('make_graph', env.make_graph)
env._masks = timed('masks', env._masks)
algo.actor.forward = timed('actor', algo.actor.forward)
algo.buffer.append_batch = timed('append', algo.buffer.append_batch)
import gc  # noqa: E402
if os.environ.get('PROBE_GC_FREEZE', '1') == '1':
    gc.collect()
    gc.freeze()
    print('gc frozen; thresholds', gc.get_threshold(), 'counts', gc.get_count())
for i in range(16):
    phase.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = vr.step(prob=0.5)
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) * 1e3
    print('step', i, 'edges', r['edge_count'], 'ms %.2f' % tot, {k: round(v, 2) for k, v in phase.items()})

# which call blocks?  trace every Python / C call of a few steps and report the leaf-most ones that took > 5 ms
stack, slow = [], []


def tracer(frame, event, arg):
    if event in ('call', 'c_call'):
        name = arg.__qualname__ if event == 'c_call' and hasattr(arg, '__qualname__') else (getattr(arg, '__name__', None) or frame.f_code.co_name)
        stack.append((name, frame.f_code.co_filename.split('/')[-1], frame.f_lineno, time.perf_counter(), [False]))
    elif event in ('return', 'c_return', 'c_exception') and stack:
        name, fn, ln, t0, child_slow = stack.pop()
        dt = (time.perf_counter() - t0) * 1e3
        if dt > 5.0:
            if not child_slow[0]:
                slow.append((round(dt, 1), name, fn, ln))
            if stack:
                stack[-1][4][0] = True


for i in range(12):
    del slow[:]
    del stack[:]
    t0 = time.perf_counter()
    sys.setprofile(tracer)
    r = vr.step(prob=0.5)
    sys.setprofile(None)
    torch.cuda.synchronize()
    print('traced step', i, 'ms %.1f' % ((time.perf_counter() - t0) * 1e3), slow[:6])
