"""Wavetable oscillator parameters (reference modems_codecs/nco.py:14-32).  The table is built here with
math.sin, exactly as the reference does, and uploaded: the device never evaluates sin() itself."""
from math import pi, sin

import numpy as np


class NCO:
	def __init__(self, **kwargs):
		self.sample_rate = kwargs.get('sample_rate', 8000.0)
		self.amplitude = kwargs.get('amplitude', 10000.0)
		self.set_frequency = kwargs.get('set_frequency', 1500.0)
		self.wavetable_size = kwargs.get('wavetable_size', 256)
		self.control = 0.0
		self.phase_accumulator = 0.0
		self.wavetable = [self.amplitude * sin(i * 2.0 * pi / self.wavetable_size) for i in range(self.wavetable_size)]
		self.index_scaling_factor = self.wavetable_size / (2.0 * pi)
		self.phase_scaling_factor = 2.0 * pi / self.sample_rate

	def describe(self, loop, keep):
		table = np.array(self.wavetable, dtype=np.float64)
		keep.append(table)
		loop.nco_phase_scale = self.phase_scaling_factor
		loop.nco_index_scale = self.index_scaling_factor
		loop.nco_set_frequency = float(self.set_frequency)
		loop.nco_two_pi = 2.0 * pi
		loop.nco_quarter = self.wavetable_size / 4.0
		loop.nco_wavetable = table.ctypes.data_as(type(loop.nco_wavetable))
		loop.nco_size = self.wavetable_size
