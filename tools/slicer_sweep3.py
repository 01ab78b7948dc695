"""Round 2, second sweep (one GPU, the bench hour): slicer geometry with the float64 crossing-by-crossing warm-up
(option warmup_far_f64) against the FP32 one.  Columns: slicer ms (segments + verify passes), whole step ms, repairs,
and whether the records equal those of the first line."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
stack = [chain_builder.build_chain(48000, l) for l in lines]
ref = None
print(f"{'segment':>8s} {'warm-up':>8s} {'exact':>6s} far {'slicer ms':>10s} {'total ms':>9s} {'repairs':>8s} {'segments':>9s} same")
grid = [(24576, 49152, 16384, 0), (24576, 49152, 16384, 1), (24576, 49152, 4096, 1), (24576, 32768, 4096, 1),
	(16384, 49152, 4096, 1), (16384, 32768, 4096, 1), (16384, 32768, 3072, 1),
	(12288, 49152, 4096, 1), (12288, 32768, 4096, 1), (12288, 32768, 3072, 1), (12288, 32768, 2048, 1), (12288, 24576, 4096, 1),
	(8192, 49152, 4096, 1), (8192, 32768, 4096, 1), (8192, 32768, 3072, 1), (8192, 32768, 2048, 1), (8192, 24576, 4096, 1),
	(6144, 32768, 4096, 1), (6144, 32768, 3072, 1), (4096, 32768, 4096, 1), (4096, 32768, 2048, 1)]
for seg, warm, exact, f64 in grid:
	eng = Engine(stack, segment_len=seg, warmup_len=warm, warmup_exact_len=exact, warmup_far_f64=f64)
	best = None
	for i in range(6):
		eng.run_device_ptr(dev.data_ptr(), len(audio))
		st = eng.stats()
		if i >= 2 and (best is None or st['slicer_ms'] < best['slicer_ms']):
			best = st
	recs, arena = eng.fetch()
	sig = (recs.tobytes(), arena.tobytes())
	if ref is None:
		ref = sig
	print(f"{seg:8d} {warm:8d} {exact:6d} {'f64' if f64 else 'f32'} {best['slicer_ms']:10.3f} {best['total_ms']:9.3f} {best['slicer_repairs']:8d} {best['slicer_segments']:9d} {sig == ref}", flush=True)
	eng.close()
