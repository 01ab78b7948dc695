"""Multi-GPU result check (torchrun, N ranks): ONE recording of N x `seconds` (the bench hour repeated, as bench.py
builds it) decoded through the shard link must give, on every rank, exactly the records an unsharded engine produces
for the whole recording on one GPU."""
import hashlib, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from pymodem_b200.sharded import LinkedRun, TorchExchange, plan_shards

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 1800.0
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
hour = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
	seed=1000, noise_seed=1001)[0]
n_total = len(hour) * world
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]
plans = plan_shards(n_total, world, trim_max=305, samples_per_symbol=40.0)
plan = plans[rank]
idx = np.arange(plan['audio_begin'], plan['audio_end'], dtype=np.int64) % len(hour)
local_audio = np.ascontiguousarray(hour[idx])
eng = Engine(stack, device=local)
ex = TorchExchange(torch.device("cuda", local))
link = LinkedRun(eng, rank, world, max(p['audio_end'] - p['audio_begin'] for p in plans), ex, ex.var)
recs, arena = link.run(plan, local_audio.ctypes.data, len(local_audio))
digest = hashlib.sha256(recs.tobytes() + arena.tobytes()).hexdigest()[:16]
all_digests = [None] * world
dist.all_gather_object(all_digests, (digest, len(recs), link.fallbacks))
if rank == 0:
	whole = np.ascontiguousarray(hour[np.arange(n_total, dtype=np.int64) % len(hour)])
	ref = Engine(stack, device=local)
	r2, a2 = ref.run_raw(whole)
	def rows(r, a):      # record order and every field except the arena placement (the merged arena is rank-major)
		raw = a.tobytes()
		return [(int(x['chain']), int(x['streamaddress']), raw[int(x['offset']):int(x['offset']) + int(x['len'])],
			int(x['bytes_corrected']), int(x['calculated_crc']), int(x['carried_crc']), int(x['valid_crc']), int(x['valid_header']))
			for x in r]
	ok = rows(recs, arena) == rows(r2, a2)
	print(f"{world} ranks, {n_total / 48000:g} s recording: linked result on every rank {all_digests}; "
		f"unsharded single-GPU run {len(r2)} records; identical: {ok}", flush=True)
	ref.close()
eng.close()
dist.destroy_process_group()
