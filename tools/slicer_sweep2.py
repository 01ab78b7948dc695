"""slicer geometry sweep on the bench workload (one GPU): slicer_ms per (segment_len, warmup_len)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
lines = configs.afsk_1200_ax25_super_opt()
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(lines)]
audio = synth.afsk1200_ax25(duration_s=3600.0, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
	seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda(); torch.cuda.synchronize()
n = len(audio)
for seg, warm, exact in [(24576, 49152, 16384), (32768, 49152, 16384), (28672, 49152, 16384), (20480, 49152, 16384), (24576, 65536, 16384),
		(32768, 65536, 14336), (24576, 49152, 14336), (28672, 57344, 14336), (36864, 49152, 16384)]:
	eng = Engine(stack, segment_len=seg, warmup_len=warm, warmup_exact_len=exact)
	for _ in range(3):
		eng.run_device_ptr(dev.data_ptr(), n)
	best = None
	for _ in range(4):
		eng.run_device_ptr(dev.data_ptr(), n)
		st = eng.stats()
		if best is None or st['slicer_ms'] < best['slicer_ms']:
			best = st
	print(f"seg {seg:6d} warm {warm:6d} exact {exact:6d}: slicer_ms {best['slicer_ms']:.3f} total {best['total_ms']:.3f} repairs {best['slicer_repairs']} segments {best['slicer_segments']}", flush=True)
	eng.close()
