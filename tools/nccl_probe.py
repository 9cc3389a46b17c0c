"""2+ GPU diagnostic: time an all-reduce of the flat gradient buffer size (eager and graph-captured)."""
import os, sys, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 6_270_000
g = torch.randn(n, device="cuda")
def timeit(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3
t_eager = timeit(lambda: dist.all_reduce(g))
gr = torch.cuda.CUDAGraph()
dist.all_reduce(g); torch.cuda.synchronize()
with torch.cuda.graph(gr):
    dist.all_reduce(g)
t_graph = timeit(gr.replay)
if dist.get_rank() == 0:
    print(f"all_reduce {n*4/1e6:.1f} MB world={dist.get_world_size()}: eager {t_eager:.1f} us, graph {t_graph:.1f} us", flush=True)
dist.barrier(); torch.cuda.synchronize(); os._exit(0)
