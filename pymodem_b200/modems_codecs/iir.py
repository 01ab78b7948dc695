"""First-order bilinear low-pass coefficients (reference modems_codecs/iir.py:8-36)."""
from math import pi, tan


class IIR_1:
	def __init__(self, **kwargs):
		self.sample_rate = kwargs.get('sample_rate', 8000.0)
		self.filter_type = kwargs.get('filter_type', 'lpf')
		self.cutoff_freq = kwargs.get('cutoff', 100.0)
		self.gain = kwargs.get('gain', 2.0)
		radian_cutoff = 2.0 * pi * self.cutoff_freq
		if self.filter_type != 'lpf':
			raise NotImplementedError("IIR_1 only defines the 'lpf' type (iir.py:17)")
		warp_cutoff = 2.0 * self.sample_rate * tan(radian_cutoff / (2.0 * self.sample_rate))
		omega_T = warp_cutoff / self.sample_rate
		a1 = (2.0 - omega_T) / (2.0 + omega_T)
		b0 = omega_T / (2.0 + omega_T)
		b1 = b0
		self.b_coefs = [self.gain * b0, self.gain * b1]
		self.a_coefs = [0.0, a1]
		self.order = 1

	def describe(self, loop):
		loop.iir_b0, loop.iir_b1 = self.b_coefs
		loop.iir_a1 = self.a_coefs[1]
