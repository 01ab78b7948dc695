"""QPSK phase-error lookup table (reference modems_codecs/phase_detector.py:12-45); the lookup itself
(get_qpsk_angle_error, :124-149) runs inside the device loop."""
from math import atan2, pi, sqrt

import numpy as np


class PhaseDetector:
	def __init__(self, constellation_id, granularity, gain):
		self.min_mag = granularity * .15
		self.max_mag = granularity * .76
		self.constellation_id = constellation_id
		self.granularity = granularity
		self.gain = gain
		self.qpsk_error_table = []
		for real in range(granularity):
			row = []
			for imag in range(granularity):
				mag = sqrt((real**2) + (imag**2))
				if mag >= self.min_mag and mag <= self.max_mag:
					row.append(round(gain * ((atan2(imag, real) * 180 / pi) - 45)))
				else:
					row.append(0)
			self.qpsk_error_table.append(row)

	def describe(self, loop, keep):
		table = np.array(self.qpsk_error_table, dtype=np.int32)
		keep.append(table)
		loop.pd_table = table.ctypes.data_as(type(loop.pd_table))
		loop.pd_granularity = self.granularity
