"""Probe the tcgen05 fp32 accumulation: inputs exactly representable in tf32 (lo = 0), so every error is accumulation."""
import sys, os, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200')]
from gcbf_b200 import ops, _C
dev = torch.device('cuda:0')
def tf32(t): return (t.view(torch.int32) & -8192).view(torch.float32)
for K in (64, 256, 1024, 2048, 8192):
    for mode in ('random', 'positive'):
        g = torch.Generator().manual_seed(K)
        x = torch.randn(512, K, generator=g); W = torch.randn(256, K, generator=g)
        if mode == 'positive': x, W = x.abs(), W.abs()
        x, W = tf32(x).to(dev), tf32(W).to(dev)
        ref = x.double() @ W.double().t()
        ops.GEMM_IMPL = 2; y2 = ops.linear_fwd(x, W, None, None, 0)
        ops.GEMM_IMPL = 1; y1 = ops.linear_fwd(x, W, None, None, 0)
        d2 = (y2.double() - ref); d1 = (y1.double() - ref)
        scale = ref.abs().max().item()
        print(f'K={K:5d} {mode:8s} tc: max {d2.abs().max().item()/scale:.2e} mean-signed {(d2*ref.sign()).mean().item()/scale:+.2e} | simt: max {d1.abs().max().item()/scale:.2e} mean-signed {(d1*ref.sign()).mean().item()/scale:+.2e}')
