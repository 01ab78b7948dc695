"""Hann-windowed Hilbert transformer taps (reference modems_codecs/hilbert.py:9-34)."""
from math import pi, sin

import numpy as np


class Hilbert:
	def __init__(self, **kwargs):
		self.tap_count = kwargs.get('tap_count', 21)
		self.amplitude = kwargs.get('amplitude', 1.0)
		self.window = kwargs.get('window', 'hann')
		if self.window != 'hann':
			raise NotImplementedError("hilbert.py only defines the hann window")
		self.delay = self.tap_count // 2
		self.taps = [2 / (pi * n) if n % 2 else 0 for n in range(-self.delay, -self.delay + self.tap_count)]
		N = self.tap_count - 1
		self.window_taps = [sin(pi * n / N)**2 for n in range(self.tap_count)]
		self.taps = [t * w for t, w in zip(self.taps, self.window_taps)]
		self.delay_taps = [0] * (self.delay + 1)
		self.delay_taps[0] = 1

	def describe(self, loop, keep):
		taps = np.array(self.taps, dtype=np.float64)
		keep.append(taps)
		loop.hilbert = taps.ctypes.data_as(type(loop.hilbert))
		loop.n_hilbert = self.tap_count
		loop.hilbert_delay = self.delay
