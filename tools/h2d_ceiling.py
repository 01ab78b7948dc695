"""Plain host-to-device copy ceiling of the box with 1 / 2 / 4 / ... concurrent ranks (torchrun, one process per GPU):
every active rank copies its own pinned 345.6 MB buffer (one hour of 48 kHz int16 audio) to its GPU, all at once,
timed with CUDA events after a barrier.  This is what bounds bench.py's e2e figure when the kernels are faster than
the copy: e2e.h2d_gbs of a rank cannot exceed its share here.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29556 tools/h2d_ceiling.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from pymodem_b200.sharded import bind_to_gpu_numa_node

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
numa = bind_to_gpu_numa_node(local)
if world > 1:
	dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 172_800_000
host = torch.empty(n, dtype=torch.int16).pin_memory()
host.random_(-2000, 2000)
dev = torch.empty(n, dtype=torch.int16, device="cuda")
steps = [k for k in (1, 2, 4, 8, 16) if k <= world]
for active in steps:
	best = None
	for rep in range(6):
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		if rank < active:
			dev.copy_(host, non_blocking=True)
		e1.record()
		torch.cuda.synchronize()
		ms = e0.elapsed_time(e1)
		if rep:                                   # the first repetition warms up
			best = ms if best is None else min(best, ms)
	t = torch.tensor([best if rank < active else 0.0], dtype=torch.float64, device="cuda")
	if world > 1:
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
	if rank == 0:
		worst = float(t[0])
		print(f"{active} concurrent rank(s): slowest rank {worst:.3f} ms for {2 * n / 1e6:.1f} MB = {2 * n / worst / 1e6:.1f} GB/s per rank, "
			f"{active * 2 * n / worst / 1e6:.1f} GB/s aggregate", flush=True)
if rank == 0:
	print("rank 0:", numa)
if world > 1:
	dist.destroy_process_group()
