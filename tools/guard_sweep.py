"""guard_eps sweep on the bench workload (one GPU, one hour x 8 chains): how far inside the default guard band do the
FP32 front end's sign errors actually lie?  For every guard_eps: flagged samples, fix-up time, and whether the slicer
byte streams and packet records still equal those of the default guard (2^-16)."""
import sys, os, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
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
base = None
for k in (16, 17, 18, 19, 20, 22, 24, 30):
	eng = Engine(stack, guard_eps=2.0 ** -k)
	for _ in range(2):
		eng.run_device_ptr(dev.data_ptr(), n)
	st = eng.stats()
	h = hashlib.sha256()
	for ci in range(len(stack)):
		data, addr = eng.stream(ci, 0)
		h.update(np.ascontiguousarray(data).tobytes()); h.update(np.ascontiguousarray(addr).tobytes())
	d = h.hexdigest()[:16]
	if base is None:
		base = d
	print(f"guard_eps 2^-{k}: flagged {st['guard_flagged']} fixup_ms {st['fixup_ms']:.3f} total_ms {st['total_ms']:.3f} packets {st['n_packets']} "
		f"streams digest {d} {'== default' if d == base else '!= default'}", flush=True)
	eng.close()
