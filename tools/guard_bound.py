#!/usr/bin/env python
"""CPU calibration of the FP32 sign guard's error model (csrc/front.cu).

Emulates the AFSK front end of one super-opt chain in FP32 (every FIR as a chain of float32 fused multiply-adds in
tap order, like the kernel's register-blocked FIRs) next to the reference's float64 formula (afsk.py:148-167) on
adversarial inputs, and reports the error of the soft value y in units of the two guard terms

    T_rel = (|L_mark| + g |L_space|)                      in-band magnitude scale (round 1's only term)
    T_abs = u * A_w * sum|h_bpf| * N_corr * sum|h_lpf| * (1 + g)    raw-input term: the band-pass rounds at the
                                                                  magnitude of the RAW samples (DC, hum, out-of-band)

with u = 2^-24 and A_w the largest |audio| of the samples an output depends on.  The guard flags a sample when
|y| < eps_rel * T_rel + c_abs * T_abs; this tool gives max |err| / T_rel and max |err| / T_abs so that eps_rel and
c_abs can be set with a stated margin.  No GPU needed: run  python tools/guard_bound.py [seconds]
With --gpu the FP32 values come from the engine itself (option keep_soft with both guard terms set to 0, so that no
sample is replaced by its float64 re-evaluation): the kernel's real arithmetic -- sliding-window correlators, packed
FFMA2 accumulation order, sqrt.approx -- instead of the emulation.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pymodem_b200 import configs, synth                      # noqa: E402
from pymodem_b200.modems_codecs import chain_builder          # noqa: E402

FS = 48000
U = 2.0 ** -24


def fir32(x32, h64):
	"""y[n] = sum_j hr[j] x[n + j] with one float32 accumulator and an FMA per tap (product exact in float64, one
	rounding of the sum to float32 -- double rounding aside, that is fmaf)."""
	hr = np.asarray(h64[::-1], dtype=np.float32)
	n = len(x32) - len(hr) + 1
	acc = np.zeros(n, dtype=np.float32)
	x64 = x32.astype(np.float64)
	for j, h in enumerate(hr):
		acc = (acc.astype(np.float64) + float(h) * x64[j:j + n]).astype(np.float32)
	return acc


def front64(m, audio):
	x = np.convolve(audio.astype(np.float64), m.input_bpf, 'valid')
	def mag(ti, tq):
		return np.sqrt(np.convolve(x, ti, 'valid') ** 2 + np.convolve(x, tq, 'valid') ** 2)
	mk = mag(m.mark_correlator_i, m.mark_correlator_q)
	sp = mag(m._space_unit_i, m._space_unit_q)
	lm = np.convolve(mk, m.output_lpf, 'valid')
	ls = np.convolve(sp, m.output_lpf, 'valid')
	return lm - m.space_gain * ls, lm, ls


def front32(m, audio):
	x = fir32(audio.astype(np.float32), m.input_bpf)
	def mag(ti, tq):
		i, q = fir32(x, ti), fir32(x, tq)
		return np.sqrt((i.astype(np.float64) ** 2 + q.astype(np.float64) ** 2)).astype(np.float32)
	mk = mag(m.mark_correlator_i, m.mark_correlator_q)
	sp = mag(m._space_unit_i, m._space_unit_q)
	lm, ls = fir32(mk, m.output_lpf), fir32(sp, m.output_lpf)
	g = np.float32(m.space_gain)
	return (lm.astype(np.float64) - float(g) * ls.astype(np.float64)).astype(np.float32)


def window_max(a, w):
	"""max |a| over [n, n + w) for every n (the samples output n depends on)"""
	from scipy.ndimage import maximum_filter1d
	m = maximum_filter1d(np.abs(a.astype(np.float64)), size=w, origin=-(w // 2), mode='constant')
	return m[:len(a) - w + 1]


def cases(seconds):
	n = int(seconds * FS)
	rng = np.random.default_rng(7)
	t = np.arange(n) / FS
	lvl = 32767 * 10 ** (-50 / 20)
	noise = rng.normal(0, lvl, n)
	yield "noise -50 dBFS alone", noise
	yield "noise -50 dBFS + DC 32000", noise + 32000
	yield "noise -50 dBFS + 0.9 FS hum 50/100 Hz", noise + 0.6 * 32767 * np.sin(2 * np.pi * 50 * t) + 0.3 * 32767 * np.sin(2 * np.pi * 100 * t)
	yield "noise -50 dBFS + 0.9 FS 5 kHz tone", noise + 0.9 * 32767 * np.sin(2 * np.pi * 5000 * t)
	sig = synth.afsk1200_ax25(duration_s=seconds, sample_rate=FS, frame_interval_s=1.0, noise_start=0.0, noise_end=1.6,
		seed=3, noise_seed=4)[0].astype(np.float64)
	yield "bench-like AFSK + AWGN ramp", sig
	yield "clipped (x4) AFSK + AWGN", sig * 4
	weak = synth.afsk1200_ax25(duration_s=seconds, sample_rate=FS, frame_interval_s=1.0, amplitude=0.003, noise_start=0.0,
		noise_end=0.5, seed=5, noise_seed=6)[0].astype(np.float64)
	yield "AFSK at 0.003 FS + DC 30000", weak + 30000


def gpu_soft(lines, audio):
	from pymodem_b200.engine import Engine
	eng = Engine([chain_builder.build_chain(FS, l) for l in lines], keep_soft=1, guard_eps=0.0, guard_abs=0.0)
	try:
		eng.run_raw(audio)
		return [eng.soft(c).astype(np.float64) for c in range(len(lines))]
	finally:
		eng.close()


def main():
	args = [a for a in sys.argv[1:] if not a.startswith("--")]
	use_gpu = "--gpu" in sys.argv
	seconds = float(args[0]) if args else 4.0
	lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
	print(f"# {'engine (GPU) FP32 values' if use_gpu else 'float32 emulation'}, {seconds:g} s per case")
	print(f"{'case':42s} {'chain':>5s} {'rms y':>10s} {'max err':>10s} {'err/T_rel':>10s} {'err/T_abs':>10s} {'flag%':>8s}")
	for name, sig in cases(seconds):
		audio = np.clip(np.round(sig), -32768, 32767).astype(np.int16)
		soft = gpu_soft(lines, audio) if use_gpu else None
		for ci in (0, 7):
			m = chain_builder.build_chain(FS, lines[ci])[1]
			y64, lm, ls = front64(m, audio)
			y32 = soft[ci][:len(y64)] if use_gpu else front32(m, audio).astype(np.float64)
			err = np.abs(y32 - y64)
			g = m.space_gain
			t_rel = np.abs(lm) + g * np.abs(ls)
			trim = len(audio) - len(y64)
			a_w = window_max(audio, trim + 1)
			t_abs = U * a_w * np.abs(m.input_bpf).sum() * len(m.mark_correlator_i) * np.abs(m.output_lpf).sum() * (1 + g)
			r_rel = float(np.max(err / np.maximum(t_rel, 1e-30)))
			r_abs = float(np.max(err / np.maximum(t_abs, 1e-30)))
			thr = 2.0 ** -18 * t_rel + 0.25 * t_abs
			flagged = float(np.mean(np.abs(y64) < thr)) * 100
			bad = int(np.sum((err >= thr) & (np.abs(y64) < err)))
			print(f"{name:42s} {ci:5d} {np.sqrt(np.mean(y64 ** 2)):10.3g} {err.max():10.3g} {r_rel:10.3g} {r_abs:10.3g} {flagged:8.4f}"
				+ (f"  UNSAFE x{bad}" if bad else ""))


if __name__ == "__main__":
	main()
