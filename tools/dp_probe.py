"""Data-parallel overhead probe (run under torchrun): collective timings + train-step time with / without the collectives."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gcbf-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch.distributed as dist
import bench
world = int(os.environ.get('WORLD_SIZE', '1')); rank = int(os.environ.get('RANK', '0')); lr = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)


def ev_time(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if world > 1:
    big = torch.zeros(24_460_000, device=dev)
    small = torch.zeros(16, device=dev, dtype=torch.float64)
    t_big = ev_time(lambda: dist.all_reduce(big), 10)
    t_small = ev_time(lambda: dist.all_reduce(small), 50)
    if rank == 0:
        print(f'all_reduce 98 MB: {t_big:.3f} ms ({big.numel() * 4 / t_big / 1e6:.0f} GB/s algbw)   tiny all_reduce: {t_small * 1e3:.1f} us', flush=True)
sb, env, algo = bench.build_case('C2', dev, rank)
data = env.graph_from_states(sb.states.to(dev))
E = int(data.edge_index.shape[1])
t_dp = ev_time(lambda: algo.train_step(data), 10)
# the same step with the collectives disabled (each rank trains alone)
from gcbf_b200.distributed import Reducer
red = algo._reducer()
red.world = 1
t_solo = ev_time(lambda: algo.train_step(data), 10)
print(f'rank {rank}/{world} E={E}: step with collectives {t_dp:.2f} ms, same rank without collectives {t_solo:.2f} ms', flush=True)
if world > 1:
    dist.destroy_process_group()
