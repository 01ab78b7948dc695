"""PI controller parameters (reference modems_codecs/pi_control.py:8-13); update_saturate runs on the device."""


class PI_control:
	def __init__(self, **kwargs):
		self.p_rate = kwargs.get('p', 0.1)
		self.i_rate = kwargs.get('i', 0.1)
		self.i_limit = kwargs.get('i_limit', 100.0)
		self.gain = kwargs.get('gain', 1000.0)
		self.integral = 0.0
		self.proportional = 0.0

	def describe(self, loop):
		loop.pi_gain = float(self.gain)
		loop.pi_p = float(self.p_rate)
		loop.pi_i = float(self.i_rate)
		loop.pi_limit = float(self.i_limit)
		loop.pi_integral0 = float(self.integral)
