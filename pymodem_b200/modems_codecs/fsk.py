"""Baseband FSK modem (reference modems_codecs/fsk.py:15-159): one input FIR,
optional negation.  Host side holds parameters/taps; the FIR runs on the GPU."""
import numpy as np
from scipy.signal import firwin

from .. import _lib
from .rrc import RRC
from .string_ops import check_boolean

# config -> (symbol_rate, filter type, lpf cutoff, span, rrc rolloff); fsk.py:25-103
_PRESETS = {
	'9600': (9600.0, 'lpf', 6000.0, 1.5, False),
	'4800': (4800.0, 'lpf', 3000.0, 1.5, False),
	'4800-rrc': (4800.0, 'rrc', None, 9, 0.2),
	'9600-rrc': (9600.0, 'rrc', None, 9, 0.2),
	'4800-gauss': (4800.0, 'lpf', 0.9 * 4800.0, 4, False),
	'9600-gauss': (9600.0, 'lpf', 0.9 * 9600.0, 4, False),
}


class FSKModem:
	modem_kind = _lib.PM_MODEM_FSK

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', '9600')
		self.sample_rate = kwargs.get('sample_rate', 96000)
		preset = _PRESETS.get(self.definition, _PRESETS['9600'])
		(self.symbol_rate, self.input_filter_type, self.input_lpf_cutoff,
			self.input_lpf_span, self.rrc_rolloff_rate) = preset
		self.invert = False
		self.tune()

	def StringOptionsRetune(self, options):      # fsk.py:110-113: only 'invert'
		self.invert = check_boolean(options.get('invert', "false"))
		self.tune()

	def tune(self):                              # fsk.py:115-147
		self.input_lpf_tap_count = round(self.sample_rate * self.input_lpf_span / self.symbol_rate)
		if self.input_filter_type == 'rrc':
			self.rrc = RRC(sample_rate=self.sample_rate, symbol_rate=self.symbol_rate,
				symbol_span=self.input_lpf_span, rolloff_rate=self.rrc_rolloff_rate)
			self.input_lpf = self.rrc.taps
		else:
			self.input_lpf = firwin(self.input_lpf_tap_count, [self.input_lpf_cutoff],
				pass_zero='lowpass', fs=self.sample_rate)
		# (the reference also constructs an AGC here, fsk.py:140-147, but never applies it)

	def describe(self, desc, keep):
		a = np.ascontiguousarray(self.input_lpf, dtype=np.float64)
		keep.append(a)
		desc.modem_kind = self.modem_kind
		desc.invert_soft = 1 if self.invert else 0
		desc.bpf, desc.n_bpf = a.ctypes.data_as(_lib._dp), len(a)

	def demod(self, input_audio):
		"""fsk.py:149-159 on the GPU -> float64 ndarray."""
		from ..engine import demod_only
		return demod_only(self, input_audio)
