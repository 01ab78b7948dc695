"""Timeline of one host-buffer run of the bench hour (engine option trace=1): when each chunk has landed, when its front
end, guard fix-up and batch of slicer segments are done."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine, pinned_empty
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
a = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
audio = pinned_empty(len(a)); audio[:] = a
stack = [chain_builder.build_chain(48000, l) for l in lines]
for opts in [dict(early_tail=0), dict(early_tail=1), dict(early_tail=2, early_batches=12), dict(early_tail=2, early_batches=3)]:
	eng = Engine(stack, trace=1, **opts)
	for i in range(4):
		t0 = time.perf_counter()
		eng.run(audio)
		dt = (time.perf_counter() - t0) * 1e3
	print(opts, f"{dt:.3f} ms", {k: round(v, 3) for k, v in eng.stats().items() if k.endswith('_ms')})
	rows = {}
	for label, ms in eng.trace():
		what, i = label.split()
		rows.setdefault(int(i), {})[what] = ms
	print(" chunk   copied    front    fixup segments")
	for i in sorted(rows):
		r = rows[i]
		print(f"{i:6d} " + " ".join(f"{r[k]:8.3f}" if k in r else "        " for k in ("copied", "front", "fixup", "segments")))
	eng.close()
