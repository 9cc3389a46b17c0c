"""GPU tool: where does a tcgen05 GEMM CTA spend its time?  Cycle counters of the MMA-issuing thread:
total, waiting for operands (full barrier), waiting for a free accumulator (tmem-empty), and of one epilogue warp
waiting for the accumulator (tmem-full).   python tools/gemm_timeline.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vit_cifar_b200 as vb
from vit_cifar_b200 import ops, _lib

lib = _lib.load()
lib.vitb_debug_gemm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
bf = torch.bfloat16


def run(kind, M, N, K, mode, extra="", flags=0):
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    a = torch.randn(M, K, device="cuda").to(bf); w = torch.randn(N, K, device="cuda").to(bf)
    bias = torch.zeros(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=bf)
    res = torch.randn(M, N, device="cuda").to(bf) if "res" in extra else None
    pre = torch.empty_like(out) if "pre" in extra else None
    def call():
        if kind == "fwd":
            ops.gemm_fwd(a, w, bias, res, out, pre, M, N, K, gelu="gelu" in extra)
        elif kind == "dgrad":   # dX[M,K] = dY[M,N] W[N,K]
            ops.gemm_dgrad(out, w, None, a, M, N, K)
        else:
            dw = torch.empty(N, K, device="cuda"); db = torch.empty(N, device="cuda")
            ops.gemm_wgrad(out, a, dw, db, M, N, K)
    lib.vitb_debug_gemm_timeline(None, mode)
    if kind == "wgrad":
        dw = torch.empty(N, K, device="cuda"); db = torch.empty(N, device="cuda")
        def call():
            ops.gemm_wgrad(out, a, dw, db, M, N, K)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    # GPU time of 10 back-to-back launches replayed from a CUDA graph (eager launches are CPU-bound at these sizes)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        call()
        with torch.cuda.graph(g, stream=side):
            for _ in range(10):
                call()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    g.replay()
    e.record(); torch.cuda.synchronize()
    us = s.elapsed_time(e) * 100
    dbg.zero_()
    lib.vitb_debug_gemm_timeline(dbg.data_ptr(), mode | (flags << 8))
    call(); torch.cuda.synchronize()
    lib.vitb_debug_gemm_timeline(None, 0)
    d = dbg.view(148, 8).double()
    if mode & 0x40:  # pair mode: epilogue warp 0's accumulator-full wait, leaders (even CTAs) vs peers (odd CTAs), per stream-0 tile
        lead, peer = d[0:144:2], d[1:144:2]
        nt = lead[:, 4].clamp_min(1)
        print(f"      pair mode epilogue wait_acc_full per tile: leader {(lead[:,5]/nt).mean().item():7.0f}  peer {(peer[:,5]/nt).mean().item():7.0f}")
    act = d[:, 4] > 0
    d = d[act]
    tiles = d[:, 4].mean().item()
    print(f"{kind:5s} M={M} N={N} K={K} {extra:10s} mode={mode} {us:7.1f} us | per tile cycles: total {d[:,0].mean().item()/tiles:7.0f} "
          f"wait_operands {d[:,1].mean().item()/tiles:7.0f} wait_acc_free {d[:,2].mean().item()/tiles:7.0f} w_load {d[:,3].mean().item():7.0f} "
          f"| epilogue wait_acc_full {d[:,5].mean().item()/tiles:7.0f} | issue {d[:,6].mean().item()/tiles:6.0f} commit {d[:,7].mean().item()/tiles:5.0f} | tiles/CTA {tiles:.1f} flags={flags}")


if __name__ == "__main__":
    M = 66560
    lib.vitb_debug_gemm_prefetch(2, 8)
    # resident-weight kernels with parts switched off (per-tile cycles of the dbg launch reflect the flags; the us column does not):
    # flags 1 = the epilogue hands the accumulator straight back (loads + MMAs only), 2 = the producer loads nothing (MMAs + epilogue only)
    # (only variants without a residual / z input operand: skipping the epilogue would leave its TMA loads unconsumed and trap)
    for kind, extra in (("fwd", ""), ("dgrad", "")):
        for flags in (0, 1, 2, 3):
            run(kind, M, 384, 384, 2, extra, flags)
