"""AGC parameters (reference modems_codecs/agc.py:8-24); peak_detect/apply run inside the sequential
loop kernels of csrc/loops.cu."""


class AGC:
	def __init__(self, **kwargs):
		self.attack_rate = kwargs.get('attack_rate', 500.0)
		self.decay_rate = kwargs.get('decay_rate', 50.0)
		self.sustain_time = kwargs.get('sustain_time', 1.0)
		self.sample_rate = kwargs.get('sample_rate', 8000.0)
		self.target_amplitude = kwargs.get('target_amplitude', 10000.0)
		self.record_envelope = kwargs.get('record_envelope', False)
		self.scaled_attack_rate = self.attack_rate / self.sample_rate
		self.scaled_decay_rate = self.decay_rate / self.sample_rate
		self.sustain_increment = 1 / self.sample_rate

	def describe(self, loop):
		loop.agc_scaled_attack = self.scaled_attack_rate
		loop.agc_scaled_decay = self.scaled_decay_rate
		loop.agc_sustain_time = self.sustain_time
		loop.agc_sustain_increment = self.sustain_increment
		loop.agc_target = self.target_amplitude
