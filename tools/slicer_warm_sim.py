"""CPU simulation of the slicer's speculative warm-up (csrc/slicer.cu): how often is the state a cold start reaches at
a segment boundary NOT bit-identical to the sequential loop's, as a function of the warm-up length W, the length X of
its exact sample-by-sample tail and the arithmetic of the crossing-by-crossing part before it (FP32 or float64)?
Sign streams: the oracle's AFSK demodulator on three noise regimes of the bench recording, chains 0, 3 and 7 of
afsk_1200_ax25_super_opt (59 / 79 / 109 samples per zero crossing).  Results (round 2, 4096-sample segments, 10539
hand-offs): FP32 far part X=16384: 0.28 % fail, X=8192: 28 %; float64 far part X=16384: 0.23 %, X=8192: 0.30 %,
X=4096: 0.40 %, X=2048: 2.5 %, X=1024: 13 %, X=0: 72 % (the closed form rounds once per run where the loop rounds
once per binade, so it is an ulp or two off and a few crossings of exact tail lose that); W=32768 X=2048: 3.4 %,
W=16384: 23 % (lock_rate 0.77 needs ~150 crossings from a cold start).
Test infrastructure: uses oracle/."""
import ctypes, os, subprocess, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from pymodem_b200 import configs, synth
from oracle import oracle

so = os.path.join(HERE, "sim", "_slicer_warm_sim.so")
if not os.path.exists(so):
	subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "sim", "slicer_warm_sim.c"), "-lm"])
lib = ctypes.CDLL(so)
lib.sim.restype = ctypes.c_long
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
signs = []
for (ns, ne, seed) in [(0.0, 0.3, 1), (0.6, 0.9, 2), (1.3, 1.6, 3)]:
	audio = synth.afsk1200_ax25(duration_s=100.0, sample_rate=48000, frame_interval_s=3.1, noise_start=ns, noise_end=ne, seed=seed, noise_seed=seed + 100)[0]
	for ci in (0, 3, 7):
		m = oracle.AFSKModem(48000, lines[ci]['modem']['config'], lines[ci]['modem']['options'])
		sg = np.ascontiguousarray((m.demod(audio.astype(np.float64)) >= 0).astype(np.uint8))
		signs.append(sg)

def run(L, W, X, mode, nc):
	f = s = 0
	aw = 0.0
	for sg in signs:
		nseg = ctypes.c_long(); avg = ctypes.c_double()
		f += lib.sim(sg.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(len(sg)), ctypes.c_long(L), ctypes.c_long(W), ctypes.c_long(X), mode, nc,
			ctypes.c_double(40.0), ctypes.c_double(0.77), ctypes.byref(nseg), ctypes.byref(avg))
		s += nseg.value; aw += avg.value
	return f, s, aw / len(signs)

print("segment warm-up exact-tail far(0=f32,1=f64) crossings(0=fixed W) -> failed / hand-offs, mean warm-up")
for (L, W, X, mode, nc) in [(4096, 49152, 16384, 0, 0), (4096, 49152, 16384, 1, 0), (4096, 49152, 8192, 0, 0), (4096, 49152, 8192, 1, 0),
		(4096, 49152, 4096, 1, 0), (4096, 49152, 2048, 1, 0), (4096, 49152, 1024, 1, 0), (4096, 49152, 0, 1, 0),
		(4096, 32768, 4096, 1, 0), (4096, 32768, 2048, 1, 0), (4096, 24576, 2048, 1, 0), (4096, 16384, 2048, 1, 0),
		(4096, 49152, 2048, 1, 300), (4096, 49152, 2048, 1, 200), (4096, 49152, 4096, 1, 200)]:
	f, s, aw = run(L, W, X, mode, nc)
	print(L, W, X, mode, nc, '->', f, '/', s, f'{f / s:.4f}', int(aw), flush=True)
