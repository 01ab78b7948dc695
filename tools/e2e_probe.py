"""Host-buffer runs of the bench hour (Engine.run on page-locked memory, records -> packet sequences included) under
different engine options: where does the time after the last byte of the copy go?"""
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
ref = None
grid = [dict(early_tail=0), dict(early_tail=1), dict(early_tail=2, early_batches=3), dict(early_tail=1, h2d_chunk=4 << 20), dict(early_tail=1, h2d_chunk=16 << 20),
	dict(early_tail=0, h2d_chunk=16 << 20)]
for opts in grid:
	eng = Engine(stack, **opts)
	ts = []
	for i in range(11):
		t0 = time.perf_counter()
		out = eng.run(audio)
		ts.append((time.perf_counter() - t0) * 1e3)
	st = eng.stats()
	sig = [len(p) for p in out]
	if ref is None:
		ref = sig
	ts = ts[3:]
	print(opts, f"mean {np.mean(ts):.3f} min {np.min(ts):.3f} ms;", {k: round(v, 3) for k, v in st.items() if k.endswith('_ms')}, 'launches', st['kernel_launches'], 'same', sig == ref, flush=True)
	eng.close()
