"""pm_engine_slice_soft on soft values with long constant stretches (no zero crossings) against the oracle's slicer, with the
exact-repetition shortcut of the repairs (engine option quiet_skip) on and off, two segment lengths, six rate/preset pairs."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
from oracle import oracle
from pymodem_b200.engine import Engine, _NoModem
from pymodem_b200.modems_codecs import slicer as slicer_mod, lfsr, ax25
for rate, config in [(48000, '1200'), (44100, '1200'), (48000, '9600'), (48000, '300'), (48000, '4800'), (22050, '1200')]:
	rng = np.random.default_rng(rate + len(config))
	parts = []
	for k in range(7):
		parts.append(rng.normal(0.0, 1.0, int(rng.integers(3000, 60000))))
		level = [1.0, -1.0, 0.0, -0.25][k % 4]
		parts.append(np.full(int(rng.integers(40000, 260000)), level))
	soft = np.concatenate(parts + [rng.normal(0.0, 1.0, 5000)])
	wb, wa = oracle.BinarySlicer(rate, config, {}).slice(soft)
	for qs in (0, 1):
		for seg in (24576, 4096):
			sl = slicer_mod.BinarySlicer(sample_rate=rate, config=config)
			eng = Engine([["x", _NoModem(), sl, lfsr.LFSR(), ax25.AX25Codec(ident="x")]], quiet_skip=qs, segment_len=seg)
			b, a = eng.slice_soft(0, soft, None)
			ok = len(b) == len(wb) and np.array_equal(b, wb) and np.array_equal(a, wa)
			first = None
			if not ok:
				n = min(len(b), len(wb))
				bad = np.nonzero((b[:n] != wb[:n]) | (a[:n] != wa[:n]))[0]
				first = (int(bad[0]), int(a[bad[0]]), int(wa[bad[0]])) if len(bad) else ('len', len(b), len(wb))
			print(rate, config, 'quiet_skip', qs, 'seg', seg, 'ok', ok, first, eng.stats()['slicer_repairs'], flush=True)
			eng.close()
