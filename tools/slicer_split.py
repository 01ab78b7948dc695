"""slicer_segments_kernel alone (kernel_times pass) under different geometries: what the crossing-by-crossing part and
the exact part of a thread's chain cost, and what the verify passes cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=3600.0, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
stack = [chain_builder.build_chain(48000, l) for l in lines]
print("segment warm-up exact far | segments ms, verify ms (launches), repairs")
for seg, warm, exact, f64 in [(24576, 49152, 4096, 1), (24576, 49152, 4096, 0), (24576, 4096, 0, 1), (24576, 0, 0, 1), (24576, 24576, 4096, 1), (24576, 98304, 4096, 1),
		(12288, 49152, 4096, 1), (12288, 4096, 0, 1), (12288, 0, 0, 1), (49152, 49152, 4096, 1), (49152, 0, 0, 1), (6144, 0, 0, 1), (6144, 49152, 4096, 1)]:
	eng = Engine(stack, segment_len=seg, warmup_len=warm, warmup_exact_len=exact, warmup_far_f64=f64, kernel_times=1)
	for _ in range(3):
		eng.run_device_ptr(dev.data_ptr(), len(audio))
	kt = {n: (c, ms) for n, c, ms in eng.kernel_times()}
	st = eng.stats()
	print(seg, warm, exact, 'f64' if f64 else 'f32', '|', round(kt['slicer_segments_kernel'][1], 3), round(kt['slicer_verify_kernel'][1], 3), f"({kt['slicer_verify_kernel'][0]})", st['slicer_repairs'], flush=True)
	eng.close()
