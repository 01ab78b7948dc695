"""Root-raised-cosine tap design (host side; reference modems_codecs/rrc.py:18-95,
default 'rect' window -- the only one any modem selects)."""
import math

import numpy as np


class RRC:
	def __init__(self, **kwargs):
		self.sample_rate = kwargs.get('sample_rate', 44100)
		self.symbol_span = kwargs.get('symbol_span', 8)
		self.symbol_rate = kwargs.get('symbol_rate', 300)
		self.rolloff_rate = kwargs.get('rolloff_rate', 0.3)
		self.window = kwargs.get('window', "rect")
		if self.window != 'rect':
			raise NotImplementedError("only the rect window is used on the demod_chain path")
		self.tune()

	def tune(self):
		beta = self.rolloff_rate
		self.oversample = self.sample_rate / self.symbol_rate
		self.tap_count = int(round(self.symbol_span * self.oversample, 0)) + 1
		self.time_step = 1 / self.sample_rate
		self.symbol_time = Ts = 1 / self.symbol_rate
		# rrc.py:23 -- float arange: the tap count is whatever it yields
		self.time = np.arange(0, self.tap_count * self.time_step, self.time_step) \
			- (self.tap_count * self.time_step / 2) + (self.time_step / 2)
		self.tap_count = len(self.time)
		try:
			asymptote = Ts / (4 * beta)
		except ZeroDivisionError:
			asymptote = False
		taps = []
		for t in self.time:
			if math.isclose(t, -asymptote) or math.isclose(t, asymptote):
				num = beta * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) + (1 - (2 / np.pi)) * np.cos(np.pi / (4 * beta)))
				taps.append(num / (Ts * pow(2, 0.5)))
			else:
				num = np.sin(np.pi * t * (1 - beta) / Ts) + 4 * beta * t * np.cos(np.pi * t * (1 + beta) / Ts) / Ts
				den = np.pi * t * (1 - pow(4 * beta * t / Ts, 2)) / Ts
				taps.append(num / (den * Ts))
		taps = taps / np.linalg.norm(taps)
		self.filter_window = [1] * self.tap_count
		self.taps = np.multiply(taps, self.filter_window)
