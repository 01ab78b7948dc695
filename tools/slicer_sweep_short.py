"""Slicer geometry for a SMALL shard (one hour split over 8 GPUs = 450 s per rank): segment / warm-up / exact-tail
lengths against slicer time and repairs.  Short segments make repairs cheap, so the warm-up can be much shorter than
the 49152 samples that are best for a whole hour on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 450.0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
stack = [chain_builder.build_chain(48000, l) for l in lines]
ref = None
print(f"{'segment':>8s} {'warm-up':>8s} {'exact':>6s} {'slicer ms':>10s} {'total ms':>9s} {'repairs':>8s} {'segments':>9s} same")
for seg, warm, exact in [(24576, 49152, 16384), (8192, 49152, 16384), (4096, 49152, 16384), (4096, 32768, 8192), (4096, 16384, 8192), (4096, 16384, 4096),
		(4096, 8192, 4096), (2048, 16384, 4096), (2048, 8192, 4096), (2048, 8192, 2048), (1024, 8192, 2048)]:
	eng = Engine(stack, segment_len=seg, warmup_len=warm, warmup_exact_len=exact)
	for _ in range(3):
		eng.run_device_ptr(dev.data_ptr(), len(audio))
	st = eng.stats()
	recs, arena = eng.fetch()
	sig = (recs.tobytes(), arena.tobytes())
	if ref is None:
		ref = sig
	print(f"{seg:8d} {warm:8d} {exact:6d} {st['slicer_ms']:10.3f} {st['total_ms']:9.3f} {st['slicer_repairs']:8d} {st['slicer_segments']:9d} {sig == ref}", flush=True)
	eng.close()
