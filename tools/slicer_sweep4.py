"""Round 2, third sweep (one GPU, the bench hour): little or no speculative warm-up.  A segment that starts cold ends in
the right state anyway once it has seen ~150 zero crossings, so the verify pass can do the warm-up's job: every segment
re-runs from its predecessor's end state until it meets one of its own checkpoints (merge), which costs about half a
segment of exact steps instead of two segments of crossing-by-crossing warm-up per segment."""
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
print(f"{'segment':>8s} {'warm-up':>8s} {'exact':>6s} {'chk':>5s} {'slicer ms':>10s} {'total ms':>9s} {'repairs':>8s} {'segments':>9s} same")
grid = [(24576, 49152, 4096, 1024),
	(24576, 0, 0, 1024), (24576, 0, 0, 512), (24576, 0, 0, 2048), (32768, 0, 0, 1024), (49152, 0, 0, 1024), (16384, 0, 0, 1024), (36864, 0, 0, 1024),
	(24576, 4096, 0, 1024), (24576, 8192, 0, 1024), (24576, 8192, 2048, 1024), (24576, 16384, 2048, 1024), (24576, 16384, 4096, 1024),
	(32768, 8192, 2048, 1024), (32768, 16384, 2048, 1024), (49152, 8192, 2048, 1024), (49152, 16384, 4096, 1024)]
for seg, warm, exact, chk in grid:
	eng = Engine(stack, segment_len=seg, warmup_len=warm, warmup_exact_len=exact, checkpoint_len=chk, warmup_far_f64=1)
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
	print(f"{seg:8d} {warm:8d} {exact:6d} {chk:5d} {best['slicer_ms']:10.3f} {best['total_ms']:9.3f} {best['slicer_repairs']:8d} {best['slicer_segments']:9d} {sig == ref}", flush=True)
	eng.close()
