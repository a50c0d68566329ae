"""Is there a periodic host-side stall on this box that has nothing to do with this library?  A bare loop of one tiny torch kernel +
synchronize, timing every iteration.  `python tools/stall_probe.py`"""
import time

import torch

x = torch.zeros(1024, device='cuda')
torch.cuda.synchronize()
ts = []
t_end = time.perf_counter() + 2.0
while time.perf_counter() < t_end:
    t0 = time.perf_counter()
    x.add_(1.0)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
ts_sorted = sorted(ts)
print('iterations', len(ts), 'median ms %.4f' % ts_sorted[len(ts) // 2], 'max ms %.2f' % ts_sorted[-1])
slow = [(i, round(t, 2)) for i, t in enumerate(ts) if t > 5.0]
print('iterations slower than 5 ms:', len(slow), slow[:20])
acc, marks = 0.0, []
for i, t in enumerate(ts):
    acc += t
    if t > 5.0:
        marks.append(round(acc, 1))
print('wall-clock position (ms) of the slow ones:', marks[:20])
