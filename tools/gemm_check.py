"""GPU check of the tcgen05 3xTF32 GEMM against fp64 and against the fp32 SIMT kernel, plus timing."""
import sys, os, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'gcbf-pytorch_b200')]
from gcbf_b200 import ops, _C
dev = torch.device('cuda:0')
print('has tcgen05:', _C.lib().gcbf_has_tcgen05())
def run(M, N, K, impl):
    ops.GEMM_IMPL = impl
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g); W = torch.randn(N, K, generator=g) / math.sqrt(K); b = torch.randn(N, generator=g)
    dz = torch.randn(M, N, generator=g); rs = torch.randn(M, K, generator=g)
    xd, Wd, bd, dzd, rsd = x.to(dev), W.to(dev), b.to(dev), dz.to(dev), rs.to(dev)
    y = ops.linear_fwd(xd, Wd, bd, None, ops.ACT_RELU)
    dx = ops.linear_bwd_data(dzd, Wd, None, rsd)
    dW, db = ops.linear_bwd_weight(dzd, xd, None)
    torch.cuda.synchronize()
    impl_used = _C.lib().gcbf_last_gemm_impl()
    x64, W64, dz64 = xd.double(), Wd.double(), dzd.double()
    ry = torch.relu(x64 @ W64.t() + bd.double()); rdx = (dz64 @ W64) * (rsd > 0); rdW = dz64.t() @ x64
    e = lambda a, r: ((a.double() - r).abs().max() / r.abs().max()).item()
    return impl_used, e(y, ry), e(dx, rdx), e(dW, rdW), e(db, dz64.sum(0))
for shape in [(256, 256, 64), (384, 128, 96), (1000, 2048, 2048), (2500, 256, 2048), (777, 2048, 260), (4096, 512, 1024), (300, 130, 100)]:
    try:
        r2 = run(*shape, 0)
        r1 = run(*shape, 1)
        print(shape, 'auto impl', r2[0], 'err y/dx/dW/db: %.2e %.2e %.2e %.2e' % r2[1:], '| simt: %.2e %.2e %.2e' % r1[1:4])
    except Exception as ex:
        print(shape, 'FAILED', ex)
# timing at the real layer size
ops.GEMM_IMPL = 0
M, N, K = 24196, 2048, 2048
x = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / 45; b = torch.zeros(N, device=dev); dz = torch.randn(M, N, device=dev)
for name, fn in [('fwd', lambda: ops.linear_fwd(x, W, b, None, ops.ACT_RELU)), ('dgrad', lambda: ops.linear_bwd_data(dz, W, None, x)),
                 ('wgrad', lambda: ops.linear_bwd_weight(dz, x, None))]:
    for impl in (0, 1):
        ops.GEMM_IMPL = impl
        for _ in range(2): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'{name} impl={impl} used={_C.lib().gcbf_last_gemm_impl()} {ms:.3f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent, incl. operand prep)')
