"""Groundwork for DESIGN.md section 8 item 1 (low-pass FIR on the tensor cores): how accurate is a 100-tap FIR whose
FP32 operands are split into bf16 pieces (exact bf16 x bf16 products, FP32 accumulation), compared with the plain FP32
FFMA FIR the front end runs today?  Pure numpy emulation, no GPU.  Error is reported relative to the guard scale
|L_mark| + g |L_space| ~ L(|m|), like the front end's sign guard (2^-18 = 3.8e-6; the FFMA path's largest error over
1.4e9 samples is just under 2^-20 = 9.5e-7)."""
import numpy as np
from scipy.signal import firwin


def bf16(x):
	"""round-to-nearest-even to bfloat16, returned as float32"""
	u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
	u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
	return u.astype(np.uint32).view(np.float32)


def split(x, pieces):
	out, rest = [], np.asarray(x, dtype=np.float32)
	for _ in range(pieces):
		p = bf16(rest)
		out.append(p)
		rest = (rest.astype(np.float64) - p.astype(np.float64)).astype(np.float32)
	return out


def fir_f32(h, m):
	"""sequential FP32 accumulation, one rounding per FMA (the FFMA kernel)"""
	n = len(m) - len(h) + 1
	acc = np.zeros(n, dtype=np.float32)
	for k in range(len(h)):
		acc = (acc.astype(np.float64) + np.float64(h[k]) * m[k:k + n].astype(np.float64)).astype(np.float32)
	return acc


def fir_split(h, m, pieces, terms):
	"""sum over the chosen (i, j) piece products, each product exact, FP32 accumulation in blocks of 16 taps"""
	hs, ms = split(h, pieces), split(m, pieces)
	n = len(m) - len(h) + 1
	acc = np.zeros(n, dtype=np.float32)
	for (i, j) in terms:
		for k0 in range(0, len(h), 16):
			part = np.zeros(n, dtype=np.float64)
			for k in range(k0, min(k0 + 16, len(h))):
				part += np.float64(hs[i][k]) * ms[j][k:k + n].astype(np.float64)
			acc = (acc.astype(np.float64) + part).astype(np.float32)     # one FP32 rounding per MMA (k = 16)
	return acc


if __name__ == "__main__":
	rng = np.random.default_rng(5)
	h = firwin(100, 1400.0, fs=48000.0).astype(np.float32)             # afsk.py '1200' output low-pass at 48 kHz
	n = 200000
	t = np.arange(n + 99)
	m = (4e5 * (1.0 + 0.6 * np.sin(2 * np.pi * t / 40.0 / 7.3)) + 3e4 * np.abs(rng.standard_normal(n + 99))).astype(np.float32)
	ref = np.convolve(m.astype(np.float64), h.astype(np.float64)[::-1], 'valid')
	scale = np.convolve(np.abs(m).astype(np.float64), np.abs(h).astype(np.float64)[::-1], 'valid')
	hr = h[::-1].copy()
	def report(name, y):
		e = np.abs(y.astype(np.float64) - ref) / scale
		print(f"{name:34s} rms {np.sqrt(np.mean(e ** 2)):.2e}  max {e.max():.2e}  (2^{np.log2(e.max()):.1f})")
	report("FP32 FFMA (today)", fir_f32(hr, m))
	all9 = [(i, j) for i in range(3) for j in range(3)]
	six = [(i, j) for (i, j) in all9 if i + j <= 2]
	report("3 x bf16, 6 largest products", fir_split(hr, m, 3, six))
	report("3 x bf16, 8 products (no 3x3)", fir_split(hr, m, 3, [p for p in all9 if p != (2, 2)]))
	report("3 x bf16, all 9 products", fir_split(hr, m, 3, all9))
	report("2 x bf16, 4 products", fir_split(hr, m, 2, [(0, 0), (0, 1), (1, 0), (1, 1)]))
