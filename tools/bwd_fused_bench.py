"""Time the one-pass backward of a Linear (vitb_gemm_bwd_fused) against dgrad + wgrad on the same operands, in isolation:
CUDA events around `reps` calls that rotate through `nbuf` operand sets (together larger than the 126 MB L2).

    python tools/bwd_fused_bench.py [M N K ...]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_cifar_b200  # noqa: E402,F401
from vit_cifar_b200 import ops  # noqa: E402


def timed(fn, reps=20, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(reps):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


def main():
    shapes = [(66560, 384, 384), (17408, 384, 384), (8320, 384, 384)]
    if len(sys.argv) > 3:
        v = [int(a) for a in sys.argv[1:]]
        shapes = [tuple(v[i:i + 3]) for i in range(0, len(v), 3)]
    for M, N, K in shapes:
        nbuf = max(2, int(400e6 // (M * (N + 3 * K) * 2)) + 1)
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda *s: torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)  # noqa: E731
        dy = [mk(M, N) for _ in range(nbuf)]; x = [mk(M, K) for _ in range(nbuf)]; z = [mk(M, K) for _ in range(nbuf)]
        dx = [torch.empty(M, K, dtype=torch.bfloat16, device="cuda") for _ in range(nbuf)]
        w = mk(N, K) * (N ** -0.5)
        dw = torch.empty(N, K, device="cuda"); cs = torch.empty(K, device="cuda")
        for with_z in (False, True):
            zz = (lambda i: z[i % nbuf]) if with_z else (lambda i: None)
            t_f = timed(lambda i: ops.gemm_bwd_fused(dy[i % nbuf], x[i % nbuf], w, zz(i), dx[i % nbuf], dw, cs if with_z else None, M, N, K))
            t_d = timed(lambda i: ops.gemm_dgrad(dy[i % nbuf], w, zz(i), dx[i % nbuf], M, N, K))
            t_w = timed(lambda i: ops.gemm_wgrad(dy[i % nbuf], x[i % nbuf], dw, cs if with_z else None, M, N, K))
            fl = 4.0 * M * N * K
            print(f"M={M} N={N} K={K} z={int(with_z)}: fused {t_f:7.1f} us ({fl / t_f / 1e6:6.0f} TFLOP/s)   dgrad {t_d:6.1f} + wgrad {t_w:6.1f} = {t_d + t_w:6.1f} us"
                  f"   ({nbuf} operand sets)", flush=True)


main()
