"""Where does the time of a sharded step go?  One GPU, the bench workload, the shard protocol driven by hand
(world = 1 and an emulated world = 2 on the same device), wall clock per C-ABI call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from pymodem_b200.sharded import plan_shards

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
lines = configs.afsk_1200_ax25_super_opt()
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(lines)]
audio = synth.afsk1200_ax25(duration_s=secs, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
	seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
torch.cuda.synchronize()
eng = Engine(stack)
n = len(audio)
for it in range(4):
	t0 = time.perf_counter(); eng.run_device_ptr(dev.data_ptr(), n); t1 = time.perf_counter()
print("unsharded run_device wall ms", (t1 - t0) * 1e3, eng.stats())
plan = plan_shards(n, 1)[0]
plan['tail_bits'] = 0
for it in range(4):
	t = [time.perf_counter()]
	st = eng.shard_begin(dev.data_ptr(), n, plan, True); t.append(time.perf_counter())
	st2, ch = eng.shard_handoff(None); t.append(time.perf_counter())
	eng.shard_gather([0] * eng.n_chains); t.append(time.perf_counter())
	eng.shard_finish(None); t.append(time.perf_counter())
print("sharded(world=1) begin/handoff/gather/finish ms", [round((b - a) * 1e3, 3) for a, b in zip(t, t[1:])], eng.stats())
