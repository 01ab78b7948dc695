"""Where the tensor-core route's soft values differ from the float64 formula (afsk.py:148-167 via numpy)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import numpy as np
from guard_bound import front64, FS
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
tensor = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=FS, frame_interval_s=1.0, noise_start=0.0, noise_end=1.6, seed=3, noise_seed=4)[0]
eng = Engine([chain_builder.build_chain(FS, l) for l in lines], keep_soft=1, guard_eps=0.0, guard_abs=0.0, tensor_lpf=tensor, tc_debug=int(sys.argv[3]) if len(sys.argv) > 3 else 0, debug_sync=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
eng.run_raw(audio)
for ci in (0, 7):
	m = chain_builder.build_chain(FS, lines[ci])[1]
	y64, lm, ls = front64(m, audio)
	soft = eng.soft(ci).astype(np.float64)[:len(y64)]
	t_rel = np.abs(lm) + m.space_gain * np.abs(ls)
	rel = np.abs(soft - y64) / np.maximum(t_rel, 1e-30)
	bad = np.nonzero(rel > 3e-6)[0]
	print(f"chain {ci}: {len(y64)} samples, rms rel err {np.sqrt(np.mean(rel ** 2)):.3e}, median {np.median(rel):.3e}, max {rel.max():.3e}, {len(bad)} above 3e-6")
	if len(bad):
		print("  mod 64:", np.bincount(bad % 64, minlength=64))
		print("  tile:", np.bincount(bad // 8192))
		print("  first:", bad[:12], "t_rel there", t_rel[bad[:6]], "rel", rel[bad[:6]])
eng.close()
