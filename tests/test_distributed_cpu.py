"""world_size-2 gloo tests (CPU) of the host-side data-parallel logic: graph sharding, the reduction of the loss
partial sums to global masked means, the h_dot gather and the single flat-bucket gradient all-reduce."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcbf_b200.distributed import Reducer, shard_range


def test_shard_range_partitions_everything():
    for n in (1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _partials(h, safe, unsafe, eps=0.02):
    """torch restatement of the first six entries of gcbf_loss_partials (sum / count / ok for unsafe and safe)."""
    p = torch.zeros(16, dtype=torch.float64)
    p[0] = torch.relu(h[unsafe] + eps).double().sum()
    p[1] = unsafe.sum()
    p[2] = (h[unsafe] < 0).sum()
    p[3] = torch.relu(-h[safe] + eps).double().sum()
    p[4] = safe.sum()
    p[5] = (h[safe] >= 0).sum()
    p[7] = h.numel()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        red = Reducer()
        assert red.world == world and red.rank == rank
        g = torch.Generator().manual_seed(0)
        B, n = 6, 5
        h = torch.randn(B * n, generator=g) * 0.05
        safe = torch.rand(B * n, generator=g) < 0.5
        unsafe = (torch.rand(B * n, generator=g) < 0.3) & ~safe
        lo, hi = shard_range(B, world, rank)
        sl = slice(lo * n, hi * n)
        # 1. partial sums: reduced local partials == partials of the whole batch -> identical masked means on all ranks
        p = red.sum_(_partials(h[sl], safe[sl], unsafe[sl]))
        assert torch.allclose(p, _partials(h, safe, unsafe), rtol=1e-12, atol=1e-12)
        # 2. h_dot gather keeps rank order
        hd = red.gather_cat(h[sl].clone())
        assert torch.equal(hd, h)
        # 2b. unequal shards (B not divisible by world): sizes travel on the host, the gather drops the padding
        B2 = 7
        lo2, hi2 = shard_range(B2, world, rank)
        h2 = torch.arange(B2 * n, dtype=torch.float32)
        sizes = red.sizes((hi2 - lo2) * n)
        assert sizes == [(b - a) * n for a, b in (shard_range(B2, world, r) for r in range(world))] and sum(sizes) == B2 * n
        assert torch.equal(red.gather_cat(h2[lo2 * n:hi2 * n].clone(), sizes), h2)
        # 3. one flat-bucket all-reduce == sum of the per-rank gradients; weights stay replicated
        from gcbf_b200.algo.gcbf import _FlatBucket
        torch.manual_seed(3)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        bucket = _FlatBucket([net[0], net[1]], torch.device('cpu'))
        x = torch.randn(8, 4, generator=g)
        xs = x[rank::world]
        bucket.zero_grad()
        net(xs).square().sum().backward()
        assert net[0].weight.grad.data_ptr() == bucket.grad.data_ptr()          # autograd accumulated in place
        red.sum_(bucket.grad)
        ref = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        ref.load_state_dict(net.state_dict())
        ref(x).square().sum().backward()
        # layout: every parameter on a 256-byte boundary (64 floats), zero padding in between, module ranges back to back
        assert bucket.offsets == [0, 64, 128, 192] and bucket.ranges == [(0, 128), (128, 256)] and bucket.grad.numel() == 256
        used = torch.zeros(256, dtype=torch.bool)
        for p_, q_, o in zip(net.parameters(), ref.parameters(), bucket.offsets):
            assert p_.data_ptr() == bucket.flat.data_ptr() + 4 * o
            assert torch.allclose(bucket.grad[o:o + q_.numel()], q_.grad.reshape(-1), rtol=1e-5, atol=1e-6)
            used[o:o + q_.numel()] = True
        assert bucket.grad[~used].abs().sum() == 0 and bucket.flat[~used].abs().sum() == 0
        torch.save(bucket.grad.clone(), os.path.join(tmp, f'g{rank}.pt'))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_reductions(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g0, g1 = torch.load(tmp_path / 'g0.pt'), torch.load(tmp_path / 'g1.pt')
    assert torch.equal(g0, g1)       # every rank ends up with the same reduced gradient
