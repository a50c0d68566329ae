"""Per-rank step timing variants (debug): value loop with/without the GEMM event timer, with/without per-step host sync."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gcbf-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch.distributed as dist
from gcbf_b200 import ops, _C
import bench
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); lr = int(os.environ.get('LOCAL_RANK', '0'))
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', lr))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
sb, env, algo = bench.build_case('C2', dev, rank)
data = env.graph_from_states(sb.states.to(dev))
def loop(n, timer, sync, rebuild):
    if timer: ops.GEMM_TIMER.enable()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter(); cpu = 0.0
    for _ in range(n):
        c0 = time.perf_counter()
        d = env.graph_from_states(sb.states.to(dev)) if rebuild else data
        r = algo.train_step(d)
        cpu += time.perf_counter() - c0
        if sync: r['scalars'].cpu()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    if timer: ops.GEMM_TIMER.disable()
    return dt, cpu / n * 1e3
for _ in range(3): algo.train_step(data)
for name, kw in [('plain', dict(timer=False, sync=False, rebuild=False)), ('timer', dict(timer=True, sync=False, rebuild=False)),
                 ('sync', dict(timer=False, sync=True, rebuild=False)), ('sync+rebuild', dict(timer=False, sync=True, rebuild=True)),
                 ('plain2', dict(timer=False, sync=False, rebuild=False))]:
    dt, cpu = loop(8, **kw)
    print(f'rank {rank}/{world} {name:14s} {dt:7.2f} ms/step   (host time inside train_step calls {cpu:6.2f} ms)', flush=True)
if world > 1: dist.destroy_process_group()
