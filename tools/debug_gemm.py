"""GPU debugging aid: run the tcgen05 GEMM variants on structured inputs and print where the result deviates.
    python tools/debug_gemm.py            (each variant in its own subprocess, so a trap does not hide the rest)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def errmap(got, ref, bm=32, bn=32, limit=12):
    import torch
    d = (got.double().cpu() - ref.double()).abs()
    M, N = d.shape
    scale = ref.double().abs().max().item() + 1e-30
    rows = []
    for i in range(0, M, bm):
        rows.append(" ".join(f"{d[i:i+bm, j:j+bn].max().item() / scale:7.1e}" for j in range(0, min(N, bn * limit), bn)))
    return "\n".join(rows[:limit])


def run(variant, M, N, K):
    import torch
    from vit_cifar_b200 import ops
    torch.manual_seed(0)
    bf = torch.bfloat16
    if variant == "fwd":
        a = torch.randn(M, K).to(bf); w = (torch.randn(N, K) / K ** 0.5).to(bf)
        ref = a.float() @ w.float().t()
        out = torch.empty(M, N, dtype=bf, device="cuda")
        ops.gemm_fwd(a.cuda(), w.cuda(), None, None, out, None, M, N, K)
    elif variant == "dgrad":
        a = torch.randn(M, N).to(bf); w = (torch.randn(N, K) / N ** 0.5).to(bf)
        ref = a.float() @ w.float()
        out = torch.empty(M, K, dtype=bf, device="cuda")
        ops.gemm_dgrad(a.cuda(), w.cuda(), None, out, M, N, K)
    else:
        dy = torch.randn(M, N).to(bf); x = torch.randn(M, K).to(bf)
        ref = dy.float().t() @ x.float()
        out = torch.empty(N, K, dtype=torch.float32, device="cuda")
        ops.gemm_wgrad(dy.cuda(), x.cuda(), out, None, M, N, K)
    torch.cuda.synchronize()
    rel = ((out.double().cpu() - ref.double()).norm() / ref.double().norm()).item()
    print(f"[{variant} M={M} N={N} K={K}] rel err {rel:.3e}  finite={bool(torch.isfinite(out.float()).all())}")
    if rel > 2e-2:
        print(errmap(out.float(), ref))


if __name__ == "__main__":
    if len(sys.argv) == 5:
        run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
        sys.exit(0)
    cases = [("fwd", 128, 128, 64), ("fwd", 128, 128, 128), ("fwd", 256, 128, 384), ("fwd", 260, 384, 384), ("fwd", 20000, 1152, 384),
             ("dgrad", 128, 64, 128), ("dgrad", 128, 128, 128), ("dgrad", 260, 384, 384), ("dgrad", 20000, 1152, 384),
             ("wgrad", 64, 128, 128), ("wgrad", 128, 128, 128), ("wgrad", 260, 384, 384), ("wgrad", 20000, 1152, 384)]
    for c in cases:
        r = subprocess.run([sys.executable, __file__, *map(str, c)], capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or f"[{c}] no output")
        if r.returncode != 0:
            print(f"[{c}] exit {r.returncode}: {r.stderr.strip()[-600:]}")
