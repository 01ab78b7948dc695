"""AFSK PLL modem -- parameters and tap design on the host; FIRs, AGC and the PLL on the GPU
(csrc/loops.cu).  Mirrors reference modems_codecs/afsk_pll.py:16-170."""
import numpy as np
from scipy.signal import firwin

from .. import _lib
from .agc import AGC
from .iir import IIR_1
from .nco import NCO
from .pi_control import PI_control
from .psk import _arr, ctypes_pointer

_KEYS = ('symbol_rate', 'input_bpf_low_cutoff', 'input_bpf_high_cutoff', 'input_bpf_span', 'output_lpf_cutoff',
	'output_lpf_span', 'sample_rate', 'carrier_freq')


class AFSKPLLModem:
	modem_kind = _lib.PM_MODEM_AFSK_PLL

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', '300')
		self.sample_rate = kwargs.get('sample_rate', 8000.0)
		if self.definition != '300':
			raise ValueError(f"AFSKPLLModem has no preset '{self.definition}' (afsk_pll.py:22)")
		self.agc_attack_rate, self.agc_sustain_time, self.agc_decay_rate = 500.0, 1.0, 50.0
		self.symbol_rate = 300.0
		self.input_bpf_low_cutoff, self.input_bpf_high_cutoff, self.input_bpf_span = 1500.0, 1900.0, 7.0
		self.carrier_freq = 1700.0
		self.output_lpf_cutoff, self.output_lpf_span = 240.0, 5
		self.max_freq_offset = 50
		self.LoopFilter = IIR_1(sample_rate=self.sample_rate, filter_type='lpf', cutoff=150.0, gain=1.0)
		pi_p = 0.6
		self.FeedbackController = PI_control(p=pi_p, i=pi_p / 6000, i_limit=self.max_freq_offset, gain=900)
		self.oscillator_amplitude = 1.0
		self.tune()

	def retune(self, **kwargs):                  # afsk_pll.py:59-68
		for key in _KEYS:
			setattr(self, key, kwargs.get(key, getattr(self, key)))
		self.tune()

	def StringOptionsRetune(self, options):      # afsk_pll.py:70-79
		for key in _KEYS:
			setattr(self, key, float(options.get(key, getattr(self, key))))
		self.tune()

	def tune(self):                              # afsk_pll.py:81-138
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / self.symbol_rate)
		self.output_lpf_tap_count = round(self.sample_rate * self.output_lpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		self.output_lpf = firwin(self.output_lpf_tap_count, self.output_lpf_cutoff, fs=self.sample_rate, scale=True)
		self.AGC = AGC(sample_rate=self.sample_rate, attack_rate=self.agc_attack_rate, sustain_time=self.agc_sustain_time,
			decay_rate=self.agc_decay_rate, target_amplitude=self.oscillator_amplitude, record_envelope=False)
		self.NCO = NCO(sample_rate=self.sample_rate, amplitude=self.oscillator_amplitude,
			set_frequency=self.carrier_freq, wavetable_size=256)
		self.output_sample_rate = self.sample_rate

	def describe(self, desc, keep):
		loop = _lib.LoopDesc()
		keep.append(loop)
		self.AGC.describe(loop)
		self.NCO.describe(loop, keep)
		self.LoopFilter.describe(loop)
		self.FeedbackController.describe(loop)
		desc.modem_kind = self.modem_kind
		desc.invert_soft = 0
		desc.bpf, desc.n_bpf = _arr(self.input_bpf, keep), len(self.input_bpf)
		desc.lpf, desc.n_lpf = _arr(self.output_lpf, keep), len(self.output_lpf)
		desc.loop = ctypes_pointer(loop)

	def demod(self, input_audio):
		"""afsk_pll.py:140-170 on the GPU -> float64 ndarray."""
		from ..engine import demod_only
		return demod_only(self, input_audio)
