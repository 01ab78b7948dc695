"""2-rank diagnosis (torchrun): does the engine slow down when another process drives another GPU, and does an
initialised NCCL communicator change it?  Prints wall/GPU times per rank for (a) no process group, (b) after NCCL
init, (c) the shard protocol with per-phase wall clocks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from pymodem_b200.sharded import ShardWorker, TorchExchange, plan_shards, run_protocol

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
secs = 3600.0
lines = configs.afsk_1200_ax25_super_opt()
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(lines)]
audio = synth.afsk1200_ax25(duration_s=secs, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
	seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
torch.cuda.synchronize()
eng = Engine(stack, device=local)
n = len(audio)

def runs(tag):
	ts = []
	for it in range(8):
		t0 = time.perf_counter(); eng.run_device_ptr(dev.data_ptr(), n); ts.append((time.perf_counter() - t0) * 1e3)
	s = eng.stats()
	print(f"[{tag}] rank {rank} wall ms {[round(x, 2) for x in ts]} gpu total {s['total_ms']:.2f} front {s['front_ms']:.2f} "
		f"slicer {s['slicer_ms']:.2f} bits {s['bits_ms']:.2f}", flush=True)

print("affinity", rank, len(os.sched_getaffinity(0)), os.environ.get("OMP_NUM_THREADS"), flush=True)
runs("a: no process group")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.barrier()
runs("b: after NCCL init, unsynchronised")
dist.barrier()
for it in range(4):
	dist.barrier()
	t0 = time.perf_counter(); eng.run_device_ptr(dev.data_ptr(), n); t1 = time.perf_counter()
	print(f"[c: barrier then run] rank {rank} wall {1e3 * (t1 - t0):.2f}", flush=True)
plan = plan_shards(n * world, world)[rank]
idx = np.arange(plan['audio_begin'], plan['audio_end'], dtype=np.int64) % n
loc = torch.from_numpy(audio[idx]).cuda(); torch.cuda.synchronize()
ex = TorchExchange(torch.device("cuda", local))
for it in range(6):
	ph = {}
	dist.barrier()
	t0 = time.perf_counter()
	run_protocol([ShardWorker(eng, plan, loc.data_ptr(), loc.numel(), on_device=True)], ex, ex.var, timing=ph)
	t1 = time.perf_counter()
	s = eng.stats()
	print(f"[d: protocol] rank {rank} wall {1e3 * (t1 - t0):.2f} phases { {k: round(v, 2) for k, v in ph.items()} } "
		f"front {s['front_ms']:.2f} fixup {s['fixup_ms']:.2f} slicer {s['slicer_ms']:.2f} bits {s['bits_ms']:.2f} repairs {s['slicer_repairs']}", flush=True)
dist.destroy_process_group()
