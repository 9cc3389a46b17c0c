"""Per-phase clocks of the fused data-parallel optimiser kernel (csrc/dp.cu) on every rank.  Under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/dp_phases.py

Every launch of dp_reduce_adam_kernel records {barrier A wait (block 0), block 0's share of the slice, whole grid: entry of the
last block to all stores issued, fence, barrier B} in SM clocks (vitb_debug_dp_phases).  Prints per-rank medians over the timed
steps, in microseconds at the SM clock nvidia-smi reports, and the step time."""
import ctypes as C
import os
import statistics
import subprocess
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vit_cifar_b200 as vb  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
B = int(os.environ.get("DP_PHASES_BATCH", "1024"))
torch.manual_seed(2045)
m = vb.ViT(3, 10, img_size=32, patch=8, num_layers=7, hidden=384, mlp_hidden=384, head=12).cuda()
eng = vb.TrainEngine(m, B, process_group=dist.group.WORLD, lr=1e-3, weight_decay=5e-5)
lib = vb.load_library()
buf = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.vitb_debug_dp_phases.argtypes = [C.c_void_p]
lib.vitb_debug_dp_phases(buf.data_ptr())  # before the graph is captured: the pointer is a kernel parameter
g = torch.Generator().manual_seed(rank)
x, y = torch.randn(B, 3, 32, 32, generator=g).cuda(), torch.randint(0, 10, (B,), generator=g).cuda()
rows, ms = [], []
for it in range(30):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); eng.step(x, y); e.record()
    torch.cuda.synchronize()
    if it >= 10:
        rows.append(buf[:5].tolist()); ms.append(s.elapsed_time(e))
try:
    mhz = float(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", str(local)], capture_output=True,
                               text=True).stdout.strip())
except Exception:
    mhz = 1900.0
med = [statistics.median(r[i] for r in rows) / mhz for i in range(5)]
out = [None] * world
dist.all_gather_object(out, (rank, med, statistics.median(ms), eng._dp_mode))
if rank == 0:
    print(f"world {world}, per-GPU batch {B}, mode {eng._dp_mode}; medians over 20 graph replays, microseconds (SM clock {mhz:.0f} MHz idle reading)")
    print("rank  barrierA_wait  block0_slice  grid_entry_to_stores  fence  barrierB   step_ms")
    for r, md, st, mode in sorted(out):
        print(f"{r:4d}  {md[0]:13.1f}  {md[1]:12.1f}  {md[2]:20.1f}  {md[3]:5.1f}  {md[4]:8.1f}   {st:7.3f}")
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
