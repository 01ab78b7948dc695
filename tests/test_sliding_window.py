"""Host-side checks behind the sliding-window tone correlators of the AFSK front end (csrc/front.cu SlideUnit):
the identity they rest on, restated in float64 numpy against the reference's formula (afsk.py:134-160), and the
engine's test for "these taps are a rotation" (pm_taps_are_rotation).  No GPU needed."""
import ctypes

import numpy as np
import pytest


def reference_magnitude(x, freq, rate, n_taps, gain=1.0):
	"""afsk.py:134-160: quadrature correlators cos/sin(w k), numpy 'valid' convolution, sqrt(i^2 + q^2)."""
	k = np.arange(n_taps) * (2.0 * np.pi * freq / rate)
	ci, cq = gain * np.cos(k), gain * np.sin(k)
	return np.sqrt(np.convolve(x, ci, 'valid') ** 2 + np.convolve(x, cq, 'valid') ** 2)


def sliding_magnitude(x, freq, rate, n_taps, unit=16):
	"""The kernel's scheme in float64: per unit of 16 outputs the window sum of x[m] e^{i w (m - base)} is formed once
	and then slid, one sample entering and one leaving per output; the phase origin is the unit's first sample."""
	w = 2.0 * np.pi * freq / rate
	table = np.exp(1j * w * np.arange(n_taps + unit))
	n_out = len(x) - n_taps + 1
	out = np.empty(n_out)
	for base in range(0, n_out, unit):
		s = np.sum(x[base:base + n_taps] * table[:n_taps])
		out[base] = abs(s)
		for r in range(min(unit, n_out - base) - 1):
			s += x[base + n_taps + r] * table[n_taps + r] - x[base + r] * table[r]
			out[base + r + 1] = abs(s)
	return out


@pytest.mark.parametrize("rate,freq,n_taps", [(48000, 1200.0, 40), (48000, 2100.0, 60), (44100, 2200.0, 37),
	(22050, 1300.0, 19), (96000, 1800.0, 120)])
def test_sliding_window_sum_equals_reference_magnitude(rate, freq, n_taps):
	rng = np.random.default_rng(7)
	n = 16 * 40 + n_taps - 1 + 5          # a ragged last unit too
	t = np.arange(n)
	x = 9000.0 * np.sin(2 * np.pi * 1200.0 / rate * t + 0.4) + 2500.0 * rng.standard_normal(n)
	want = reference_magnitude(x, freq, rate, n_taps)
	got = sliding_magnitude(x, freq, rate, n_taps)
	assert got.shape == want.shape
	assert np.max(np.abs(got - want)) <= 1e-9 * np.max(want)


def _is_rotation(lib, ti, tq):
	ti = np.ascontiguousarray(ti, dtype=np.float64)
	tq = np.ascontiguousarray(tq, dtype=np.float64)
	step = ctypes.c_double(0.0)
	dp = ctypes.POINTER(ctypes.c_double)
	ok = lib.pm_taps_are_rotation(ti.ctypes.data_as(dp), tq.ctypes.data_as(dp), len(ti), ctypes.byref(step))
	return bool(ok), step.value


def test_reference_correlator_taps_are_recognised_as_rotations(cuda_lib):
	"""The taps the host mirror builds (modems_codecs/afsk.py, same expressions as afsk.py:134-144) for every rate the
	tests use, with a correlator offset and a space gain; the step angle comes back to 1e-12."""
	from pymodem_b200.modems_codecs import afsk
	for rate in (22050, 32000, 44100, 48000, 96000):
		m = afsk.AFSKModem(sample_rate=rate, config='1200')
		m.retune(correlator_offset=7.5, space_gain=1.75)          # afsk.py:72-85
		for ti, tq, f in ((m.mark_correlator_i, m.mark_correlator_q, m.mark_freq),
				(m.space_correlator_i, m.space_correlator_q, m.space_freq)):
			ok, step = _is_rotation(cuda_lib, ti, tq)
			assert ok, (rate, f)
			assert abs(step - 2.0 * np.pi * (f + 7.5) / rate) < 1e-12


def test_other_taps_are_not_rotations(cuda_lib):
	k = np.arange(40) * (2.0 * np.pi * 1200.0 / 48000.0)
	ci, cq = np.cos(k), np.sin(k)
	assert _is_rotation(cuda_lib, ci, cq)[0]
	assert not _is_rotation(cuda_lib, ci * np.hamming(40), cq * np.hamming(40))[0]      # windowed correlator
	assert not _is_rotation(cuda_lib, ci, -cq[::-1])[0]                                  # not a common step
	bumped = cq.copy(); bumped[17] += 1e-6
	assert not _is_rotation(cuda_lib, ci, bumped)[0]                                     # one tap off by 1e-6
	assert not _is_rotation(cuda_lib, ci[:8], cq[:8])[0]                                 # shorter than a 16-output unit
	assert not _is_rotation(cuda_lib, np.zeros(40), np.zeros(40))[0]
