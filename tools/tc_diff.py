"""Sign words of the tensor-core route against the FFMA route at size: where do they differ (tile, row, CTA, turn)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
out = {}
guards = {} if len(sys.argv) > 2 and sys.argv[2] == "guarded" else dict(guard_eps=0.0, guard_abs=0.0)
for tensor in (0, 1):
	eng = Engine([chain_builder.build_chain(48000, l) for l in lines], tensor_lpf=tensor, **guards)
	try:
		eng.run_raw(audio)
		out[tensor] = [eng.signs(c) for c in range(8)]
		print("tensor", tensor, eng.stats()["n_packets"], "packets", "flagged", eng.stats()["guard_flagged"], flush=True)
	finally:
		eng.close()
for c in (0, 7):
	d = np.nonzero(out[0][c] != out[1][c])[0]
	print(f"chain {c}: {len(d)} of {len(out[0][c])} words differ")
	if len(d):
		n = d.astype(np.int64) * 32
		tile = n // 8192
		print("  tiles:", np.unique(tile)[:40], "...", len(np.unique(tile)), "distinct")
		print("  CTA (tile % 148):", np.bincount(tile % 148, minlength=148))
		print("  turn (tile // 148):", np.bincount(tile // 148))
		print("  row in tile:", np.bincount((n % 8192) // 64, minlength=128))
		print("  bits differing per word (first 20):", [bin(int(a ^ b)).count("1") for a, b in zip(out[0][c][d[:20]], out[1][c][d[:20]])])
