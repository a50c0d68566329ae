"""TEST INFRASTRUCTURE ONLY -- CPU model of the arithmetic of the tensor-core linear layers (csrc/gemm_tcgen05_f16.cu).

The product computes an fp32 GEMM as three fp16 tensor-core products of [hi | lo] companions.  This module restates that
arithmetic in plain torch so that (a) the companion format is pinned bit-for-bit (tests/test_kernels_gpu.py compares the
split kernel against `split`) and (b) the accuracy claims written in DESIGN.md section 5 are checked on the CPU, without a
GPU, against fp64 (tests/test_oracle_cpu.py::test_fp16x3_model_accuracy).  Only tests/ import it.
"""
import torch


def scale_for(amax: torch.Tensor) -> float:
    """Power of two s with amax * s in [2^14, 2^15) (1 for zero / denormal / non-finite amax); mirrors
    scale_bits_from_amax in csrc/gemm_tcgen05_f16.cu."""
    bits = amax.reshape(1).float().view(torch.int32).item()
    e = (bits >> 23) & 0xff
    if e in (0, 255):
        return 1.0
    se = min(max(127 + 14 - (e - 127), 2), 252)
    return 2.0 ** (se - 127)


def split(x: torch.Tensor):
    """(hi, lo, s): x * s = hi + lo with hi = fp16(x * s), lo = fp16(x * s - hi), both round-to-nearest-even."""
    s = scale_for(x.abs().max())
    xs = x.float() * s
    hi = xs.half()
    lo = (xs - hi.float()).half()
    return hi, lo, s


def gemm(a: torch.Tensor, b: torch.Tensor, chunk: int = 256) -> torch.Tensor:
    """a [M, K] x b [N, K]^T with the kernel's arithmetic: three products hi*hi + lo*hi + hi*lo (lo*lo dropped), exact
    products of fp16 values, K consumed in chunks whose partial sums are promoted to fp32 with round-to-nearest (the
    in-chunk accumulation is modelled in fp64 and rounded once: the tensor core's own in-chunk error is bounded by
    tools/acc_probe.py on hardware), descaled by 1 / (s_a * s_b)."""
    ah, al, sa = split(a)
    bh, bl, sb = split(b)
    ah, al, bh, bl = ah.double(), al.double(), bh.double(), bl.double()
    acc = torch.zeros(a.shape[0], b.shape[0], dtype=torch.float32)
    for k0 in range(0, a.shape[1], chunk):
        sl = slice(k0, k0 + chunk)
        part = ah[:, sl] @ bh[:, sl].T + al[:, sl] @ bh[:, sl].T + ah[:, sl] @ bl[:, sl].T
        acc = acc + part.float()
    return acc * (1.0 / sa) * (1.0 / sb)
