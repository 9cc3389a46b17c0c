"""The nine GEMM shapes of one training step (B = 1024, T = 65: 66,560 rows) through cuBLAS / ATen as PyTorch eager runs them
(bf16 operands: F.linear with bias = cuBLASLt epilogue; GELU, residual add, gelu', bias-gradient sum as separate ATen kernels —
what the reference executes) next to the hand-written kernels (same operands, fused epilogues), in isolation:
CUDA events around rotating operand sets larger than L2.

    python tools/cublas_shapes.py [rows]          -> one JSON object on the last line
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_cifar_b200  # noqa: E402,F401
from vit_cifar_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 66560
H = 384


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


g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda *s: torch.randn(*s, generator=g, device="cuda").to(torch.bfloat16)  # noqa: E731
NB = 3
rows = []


def add(name, flops, ours, eager):
    t_o, t_e = timed(ours), timed(eager)
    rows.append({"gemm": name, "ours_us": round(t_o, 1), "eager_us": round(t_e, 1), "ours_tflops": round(flops / t_o / 1e6, 1),
                 "eager_tflops": round(flops / t_e / 1e6, 1), "speedup": round(t_e / t_o, 2)})
    print(f"{name:44s} ours {t_o:7.1f} us  eager/cuBLAS {t_e:7.1f} us  x{t_e / t_o:.2f}", flush=True)


for N in (3 * H, H):
    K = H
    a = [mk(M, K) for _ in range(NB)]; res = [mk(M, N) for _ in range(NB)]; w = mk(N, K) * K ** -0.5; b32 = mk(N).float(); bb = b32.to(torch.bfloat16)
    out = [torch.empty(M, N, dtype=torch.bfloat16, device="cuda") for _ in range(NB)]; pre = [torch.empty_like(o) for o in out]
    fl = 2.0 * M * N * K
    add(f"fwd  {M}x{N}x{K} bias", fl, lambda i: ops.gemm_fwd(a[i % NB], w, b32, None, out[i % NB], None, M, N, K),
        lambda i: F.linear(a[i % NB], w, bb))
    if N == H:
        add(f"fwd  {M}x{N}x{K} bias+residual", fl, lambda i: ops.gemm_fwd(a[i % NB], w, b32, res[i % NB], out[i % NB], None, M, N, K),
            lambda i: F.linear(a[i % NB], w, bb) + res[i % NB])
        add(f"fwd  {M}x{N}x{K} bias+GELU (saves z)", fl, lambda i: ops.gemm_fwd(a[i % NB], w, b32, None, out[i % NB], pre[i % NB], M, N, K, gelu=True),
            lambda i: F.gelu(F.linear(a[i % NB], w, bb)))
        add(f"fwd  {M}x{N}x{K} bias+GELU+residual (saves z)", fl,
            lambda i: ops.gemm_fwd(a[i % NB], w, b32, res[i % NB], out[i % NB], pre[i % NB], M, N, K, gelu=True),
            lambda i: F.gelu(F.linear(a[i % NB], w, bb)) + res[i % NB])
    # backward of y = x W^T: dy (M, N), x (M, K)
    dy = [mk(M, N) for _ in range(NB)]; x = a; z = [mk(M, K) for _ in range(NB)]
    dx = [torch.empty(M, K, dtype=torch.bfloat16, device="cuda") for _ in range(NB)]
    dw = torch.empty(N, K, device="cuda"); db = torch.empty(N, device="cuda")
    add(f"dgrad {M}x{N}->{K}", fl, lambda i: ops.gemm_dgrad(dy[i % NB], w, None, dx[i % NB], M, N, K), lambda i: dy[i % NB] @ w)
    if N == H:
        zg = [t.float().requires_grad_(True) for t in z]

        def eager_dz(i):
            t = z[i % NB]
            return torch.ops.aten.gelu_backward(dy[i % NB] @ w, t)
        add(f"dgrad {M}x{N}->{K} x gelu'(z)", fl, lambda i: ops.gemm_dgrad(dy[i % NB], w, z[i % NB], dx[i % NB], M, N, K), eager_dz)
    add(f"wgrad {N}x{K} over {M} rows + bias grad", fl, lambda i: ops.gemm_wgrad(dy[i % NB], x[i % NB], dw, db, M, N, K),
        lambda i: (dy[i % NB].t() @ x[i % NB], dy[i % NB].sum(0)))
print(json.dumps({"rows": M, "what": "isolated, rotating operand sets; eager = torch bf16 (cuBLASLt + ATen elementwise)", "gemms": rows}))
