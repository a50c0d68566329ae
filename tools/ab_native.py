"""A/B in ONE process: the chain-level train step (library sequencing, ops.NATIVE = True) against the per-kernel Python sequencing,
alternating, device-timed, plus the host time each spends inside train_step.  python tools/ab_native.py [C2] [reps]"""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200'), ROOT]
import bench
from gcbf_b200 import ops
cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda', 0)
sb, env, algo = bench.build_case(cfg, dev, 0)
data = env.graph_from_states(sb.states.to(dev))


def run(native, n=10):
    ops.NATIVE = native
    for _ in range(2):
        algo.train_step(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    e0.record()
    for _ in range(n):
        t0 = time.perf_counter()
        algo.train_step(data)
        host += time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host / n * 1e3


for r in range(reps):
    for native in (True, False):
        ms, host = run(native)
        print(f'{cfg} rep {r} {"library " if native else "python  "} {ms:8.3f} ms/step   host inside train_step {host:7.3f} ms (includes the wait for the re-linked edge count)', flush=True)
# host cost with the device idle between steps: what the Python / ctypes layer itself costs
for native in (True, False):
    ops.NATIVE = native
    t = 0.0
    for _ in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        algo.train_step(data)
        t += time.perf_counter() - t0
        torch.cuda.synchronize()
    print(f'{cfg} {"library" if native else "python "} host time per step with an idle device at entry: {t / 6 * 1e3:.3f} ms', flush=True)
