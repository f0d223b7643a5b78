"""Per-rank host<->device copy bandwidth, ranks running simultaneously (torchrun).  Diagnoses e2e scaling."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = torch.empty(201326592 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device="cuda")
ho = torch.empty(67108864 // 4, dtype=torch.float32).pin_memory()
do = torch.empty_like(ho, device="cuda")
def bar():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
for name, fn, nbytes in (("h2d", lambda: d.copy_(h, non_blocking=True), h.numel() * 4), ("d2h", lambda: ho.copy_(do, non_blocking=True), ho.numel() * 4)):
    fn(); bar()
    t0 = time.time()
    for _ in range(10): fn()
    torch.cuda.synchronize()
    dt = (time.time() - t0) / 10
    print(f"rank {rank}: {name} {nbytes/dt/1e9:.1f} GB/s ({dt*1e3:.2f} ms)", flush=True)
    bar()
print(f"rank {rank}: cpu_count={os.cpu_count()} affinity={len(os.sched_getaffinity(0))} omp={os.environ.get('OMP_NUM_THREADS')}", flush=True)
if world > 1: dist.destroy_process_group()
