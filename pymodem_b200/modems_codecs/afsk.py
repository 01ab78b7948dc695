"""AFSK correlator modem -- parameters and tap design on the host, demodulation on
the GPU.  Mirrors reference modems_codecs/afsk.py:13-167 (class name, kwargs,
StringOptionsRetune keys, tune(), output_sample_rate, demod())."""
from math import ceil

import numpy as np
from scipy.signal import firwin

from .. import _lib

_PRESETS = {
	# afsk.py:19-42
	'300': dict(symbol_rate=300.0, input_bpf_low_cutoff=1500.0, input_bpf_high_cutoff=1900.0,
		input_bpf_span=7, mark_freq=1695.0, space_freq=1705.0, space_gain=1.0,
		output_lpf_cutoff=240.0, output_lpf_span=2.5, correlator_span=0.3, correlator_offset=0.0),
	# afsk.py:43-66 (also the fall-through for any other config string)
	'1200': dict(symbol_rate=1200.0, input_bpf_low_cutoff=900.0, input_bpf_high_cutoff=2500.0,
		input_bpf_span=3.7, mark_freq=1200.0, space_freq=2200.0, space_gain=1.0,
		output_lpf_cutoff=1400.0, output_lpf_span=2.5, correlator_span=1.0, correlator_offset=0.0),
}
_TUNABLE = ('symbol_rate', 'input_bpf_low_cutoff', 'input_bpf_high_cutoff', 'input_bpf_span',
	'mark_freq', 'space_freq', 'space_gain', 'output_lpf_cutoff', 'output_lpf_span',
	'correlator_span', 'correlator_offset', 'sample_rate')


class AFSKModem:
	modem_kind = _lib.PM_MODEM_AFSK

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', '1200')
		self.sample_rate = kwargs.get('sample_rate', 8000)
		for key, value in _PRESETS['300' if self.definition == '300' else '1200'].items():
			setattr(self, key, value)
		self.output_oversample = 1.0            # afsk.py:68 (resample_poly path is dead)
		self.tune()

	def retune(self, **kwargs):                  # afsk.py:72-85
		for key in _TUNABLE:
			setattr(self, key, kwargs.get(key, getattr(self, key)))
		self.tune()

	def StringOptionsRetune(self, options):      # afsk.py:87-100: option values are strings
		for key in _TUNABLE:
			setattr(self, key, float(options.get(key, getattr(self, key))))
		self.tune()

	def tune(self):                              # afsk.py:102-146
		fs = self.sample_rate
		self.input_bpf_tap_count = round(fs * self.input_bpf_span / self.symbol_rate)
		self.output_lpf_tap_count = round(fs * self.output_lpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count,
			[self.input_bpf_low_cutoff, self.input_bpf_high_cutoff], pass_zero='bandpass', fs=fs)
		self.output_lpf = firwin(self.output_lpf_tap_count, self.output_lpf_cutoff, fs=fs)
		time_indices = np.arange(ceil(self.correlator_span * fs / self.symbol_rate))
		mark_indices = time_indices * (2.0 * np.pi * (self.mark_freq + self.correlator_offset) / fs)
		self.mark_correlator_i = np.cos(mark_indices)
		self.mark_correlator_q = np.sin(mark_indices)
		space_indices = time_indices * (2.0 * np.pi * (self.space_freq + self.correlator_offset) / fs)
		# unit-gain tone tables are kept so chains differing only in space_gain
		# can share one correlator pass on the device
		self._space_unit_i = np.cos(space_indices)
		self._space_unit_q = np.sin(space_indices)
		self.space_correlator_i = self.space_gain * self._space_unit_i
		self.space_correlator_q = self.space_gain * self._space_unit_q
		self.output_sample_rate = self.output_oversample * self.sample_rate

	def describe(self, desc, keep):
		"""Fill the modem part of a pm_chain_desc; `keep` collects the arrays
		whose memory the descriptor points into."""
		def arr(a):
			a = np.ascontiguousarray(a, dtype=np.float64)
			keep.append(a)
			return a.ctypes.data_as(_lib._dp)
		desc.modem_kind = self.modem_kind
		desc.invert_soft = 0
		desc.bpf, desc.n_bpf = arr(self.input_bpf), len(self.input_bpf)
		desc.n_corr = len(self.mark_correlator_i)
		desc.mark_i, desc.mark_q = arr(self.mark_correlator_i), arr(self.mark_correlator_q)
		desc.space_i, desc.space_q = arr(self.space_correlator_i), arr(self.space_correlator_q)
		desc.space_unit_i, desc.space_unit_q = arr(self._space_unit_i), arr(self._space_unit_q)
		desc.space_gain = float(self.space_gain)
		desc.lpf, desc.n_lpf = arr(self.output_lpf), len(self.output_lpf)

	def demod(self, input_audio):
		"""afsk.py:148-167 on the GPU -> float64 ndarray of soft values."""
		from ..engine import demod_only
		return demod_only(self, input_audio)
