"""Per-stage SM cycles of the AFSK front-end kernel on the bench workload (option "stage_clocks"): where a CTA's time goes."""
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
for opts in ({}, {"slide_correlator": 0}):
	eng = Engine(stack, stage_clocks=1, **opts)
	for _ in range(3):
		eng.run_device_ptr(dev.data_ptr(), n)
	eng.stage_clocks()
	eng.run_device_ptr(dev.data_ptr(), n)
	st = eng.stats()
	c = eng.stage_clocks()
	ctas = c[4]
	names = ["staging", "band-pass", "correlators", "low-pass+epilogue"]
	tot = sum(c[:4])
	print(opts, f"front_ms {st['front_ms']:.3f}  CTAs {ctas}  cycles/CTA {tot / ctas:.0f}")
	for i, nm in enumerate(names):
		print(f"   {nm:18s} {c[i] / ctas:9.0f} cycles/CTA  {100 * c[i] / tot:5.1f} %")
	eng.close()
