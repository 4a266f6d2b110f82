"""Validate and time the tcgen05 3xTF32 GEMM against torch (run on the GPU box)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dreamgnn_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.double() - b).norm() / b.norm())


def timeit(fn, reps=5):
    fn(); th.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); th.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


dev = th.device('cuda:0')
g = th.Generator(dev).manual_seed(0)
shapes = [(128, 128, 32, 1), (128, 128, 64, 1), (256, 128, 128, 1), (130, 70, 36, 1), (1000, 344, 1024, 2), (257, 129, 100, 3),
          (1024, 344, 5000, 2)]
for (M, N, K, R) in shapes:
    a = th.randn(M, K, device=dev, generator=g)
    b = th.randn(R, N, K, device=dev, generator=g) if R > 1 else th.randn(N, K, device=dev, generator=g)
    ref = a.double() @ b.double().transpose(-1, -2)
    for prec in (0, 1):
        c = ops.gemm_nt(a, b, precision=prec)
        th.cuda.synchronize()
        print('M=%d N=%d K=%d R=%d prec=%d rel err %.3e (torch fp32 %.3e)' % (M, N, K, R, prec, rel(c, ref),
              rel(a @ b.transpose(-1, -2), ref)), flush=True)
for (M, N, K, R) in [(128, 128, 32, 1), (130, 70, 36, 1), (257, 129, 100, 3), (1024, 344, 5000, 2)]:
    a = th.randn(M, K, device=dev, generator=g)
    b = th.randn(R, N, K, device=dev, generator=g) if R > 1 else th.randn(N, K, device=dev, generator=g)
    ref = a.double() @ b.double().transpose(-1, -2)
    for ta, tb in ((True, False), (False, True), (True, True)):
        c = ops.gemm(a.t().contiguous() if ta else a, b.transpose(-1, -2).contiguous() if tb else b, trans_a=ta, trans_b=tb)
        th.cuda.synchronize()
        print('M=%d N=%d K=%d R=%d trans_a=%d trans_b=%d rel err %.3e' % (M, N, K, R, ta, tb, rel(c, ref)), flush=True)
if len(sys.argv) > 1:
    for (M, N, K, R) in [(1024, 344, 100000, 2), (1024, 768, 100000, 1)]:      # weight gradients: X^T dY as stored
        x = th.randn(K, M, device=dev, generator=g)
        dy = th.randn(R, K, N, device=dev, generator=g)
        t0 = timeit(lambda: ops.gemm(x, dy, trans_a=True, trans_b=True))
        t1 = timeit(lambda: ops.gemm_nt(x.t(), dy.transpose(1, 2)))
        print('dW M=%d N=%d K=%d R=%d: MN-major operands %.3f ms, with transposed copies %.3f ms' % (M, N, K, R, t0, t1), flush=True)
    for (M, N, K, R) in [(100000, 344, 1024, 2), (50000, 344, 768, 2), (100000, 768, 1024, 1), (1024, 344, 100000, 2),
                         (1024, 768, 100000, 1)]:
        a = th.randn(M, K, device=dev, generator=g)
        b = th.randn(R, N, K, device=dev, generator=g)
        t0 = timeit(lambda: ops.gemm_nt(a, b, precision=0))
        t1 = timeit(lambda: ops.gemm_nt(a, b, precision=1))
        bt = b.transpose(1, 2).contiguous()
        t2 = timeit(lambda: th.matmul(a.unsqueeze(0), bt))
        fl = 2.0 * M * N * K * R
        print('M=%d N=%d K=%d R=%d: 3xTF32 %.3f ms (%.0f TF/s eff), TF32 %.3f ms, cuBLAS fp32 %.3f ms (%.0f TF/s)'
              % (M, N, K, R, t0, fl / t0 / 1e9, t1, t2, fl / t2 / 1e9), flush=True)
