"""afsk_1200.json as shipped (two AX.25 and two IL2P chains) on ONE recording split over the ranks of a torchrun launch:
result on every rank against an unsharded single-GPU run, and the time per linked run (device-resident shards, max over
ranks).  The IL2P decoder state crosses the rank boundaries inside the link buffer (csrc/link.cu link_il2p_*).
usage: torchrun --nproc-per-node N tools/link_il2p_timing.py [seconds]   (also runs with plain python: one rank)"""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from pymodem_b200 import synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from pymodem_b200.sharded import LinkedRun, TorchExchange, local_exchange, plan_shards
from bench import records_digest
from util import Golden

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 900.0
torch.cuda.set_device(local)
if world > 1:
	dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lines = Golden("afsk1200_il2p_48k").chain_lines()
audio = synth.afsk1200_il2p(sample_rate=48000, duration_s=seconds, frame_interval_s=0.8, noise_start=0.1, noise_end=1.0, seed=85,
	noise_seed=86, first_frame_s=0.3, payload_len=[None, 300, 10, 0, 240, 60])[0]
stack = [chain_builder.build_chain(48000, l) for l in lines]
plans = plan_shards(len(audio), world, trim_max=305, samples_per_symbol=40.0)
plan = plans[rank]
dev = torch.from_numpy(np.ascontiguousarray(audio[plan['audio_begin']:plan['audio_end']])).cuda()
eng = Engine(stack, device=local)
if world > 1:
	ex = TorchExchange(torch.device("cuda", local))
	link = LinkedRun(eng, rank, world, max(p['audio_end'] - p['audio_begin'] for p in plans), ex, ex.var)
else:
	link = LinkedRun(eng, 0, 1, len(audio), local_exchange)
times = []
for rep in range(8):
	if world > 1:
		dist.barrier()
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	recs, arena = link.run(plan, dev.data_ptr(), dev.numel(), on_device=True)
	torch.cuda.synchronize()
	times.append((time.perf_counter() - t0) * 1e3)
mine = (records_digest(recs, arena), int(len(recs)), link.fallbacks, link.recoveries, min(times[2:]), eng.stats()['bits_ms'])
everyone = [mine]
if world > 1:
	everyone = [None] * world
	dist.all_gather_object(everyone, mine)
if rank == 0:
	solo = Engine(stack, device=local)
	whole = torch.from_numpy(audio).cuda()
	for _ in range(3):
		solo.run_device_ptr(whole.data_ptr(), len(audio))
	st = solo.stats()
	r2, a2 = solo.fetch()
	ok = all(e[0] == records_digest(r2, a2) and e[1] == len(r2) for e in everyone)
	print(f"afsk_1200.json (2 AX.25 + 2 IL2P chains), {seconds:g} s of 48 kHz audio, {world} rank(s): linked run "
		f"{max(e[4] for e in everyone):.3f} ms (max over ranks, best of 6; bit level per rank {[round(e[5], 3) for e in everyone]} ms), "
		f"fallbacks {[e[2] for e in everyone]}, recoveries {[e[3] for e in everyone]}; unsharded single-GPU run {st['total_ms']:.3f} ms "
		f"(bit level {st['bits_ms']:.3f}), {len(r2)} records; every rank identical to it: {ok}", flush=True)
	solo.close()
eng.close()
if world > 1:
	dist.destroy_process_group()
