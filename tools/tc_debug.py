"""Small run of the tensor-core low-pass route (for compute-sanitizer / quick checks): packets against the FFMA route."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
mask = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=0.7, noise_start=0.0, noise_end=0.8,
	seed=91, noise_seed=92, first_frame_s=0.05)[0]
res = {}
for tensor in (1, 0):
	eng = Engine([chain_builder.build_chain(48000, l) for l in lines], tensor_lpf=tensor, keep_soft=1, debug_sync=1, tc_debug=mask)
	try:
		pk = eng.run(audio)
		res[tensor] = ([[(p.streamaddress, bytes(p.data)) for p in c] for c in pk], [eng.soft(c) for c in range(8)], eng.stats())
	finally:
		eng.close()
	print("tensor_lpf", tensor, "packets", [len(c) for c in res[tensor][0]], "flagged", res[tensor][2]["guard_flagged"], flush=True)
print("packets equal:", res[0][0] == res[1][0])
for c in range(8):
	a, b = res[1][1][c].astype(np.float64), res[0][1][c].astype(np.float64)
	print(c, "soft max |tc - ffma| / rms", float(np.max(np.abs(a - b)) / np.sqrt(np.mean(b * b))))
a, b = res[1][1][7].astype(np.float64), res[0][1][7].astype(np.float64)
err = np.abs(a - b) / np.sqrt(np.mean(b * b))
bad = np.nonzero(err > 1e-5)[0]
print("chain 7: samples", len(err), "bad", len(bad), "first", bad[:10], "last", bad[-5:])
if len(bad):
	print("bad mod 8192 histogram (by row of 64):", np.bincount((bad % 8192) // 64, minlength=128))
	print("bad mod 64 histogram:", np.bincount(bad % 64, minlength=64))
	print("bad by tile:", np.bincount(bad // 8192))
	i = int(bad[len(bad) // 2])
	print("example", i, a[i - 2:i + 3], b[i - 2:i + 3])
