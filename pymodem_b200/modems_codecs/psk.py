"""PSK modems -- parameters, tap design and loop constants on the host; the FIRs, the AGC and the carrier
loops run on the GPU (csrc/loops.cu).  Mirrors reference modems_codecs/psk.py: BPSKModem (20-195) and
MPSKModem (479-773) with their presets, StringOptionsRetune keys, tune() and output_sample_rate.
(QPSKModem, type 'qpsk', is used by no shipped config and is not provided.)"""
import numpy as np
from scipy.signal import firwin

from .. import _lib
from .agc import AGC
from .hilbert import Hilbert
from .iir import IIR_1
from .nco import NCO
from .phase_detector import PhaseDetector
from .pi_control import PI_control
from .rrc import RRC


def _arr(a, keep):
	a = np.ascontiguousarray(a, dtype=np.float64)
	keep.append(a)
	return a.ctypes.data_as(_lib._dp)


class BPSKModem:
	modem_kind = _lib.PM_MODEM_BPSK

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', '300')
		self.sample_rate = kwargs.get('sample_rate', 8000.0)
		self.agc_attack_rate, self.agc_sustain_time, self.agc_decay_rate = 500.0, 1.0, 50.0
		if self.definition == '1200':            # psk.py:56-88
			self.symbol_rate = 1200.0
			self.input_bpf_low_cutoff, self.input_bpf_high_cutoff, self.input_bpf_span = 200.0, 2800.0, 4.80
			self.carrier_freq = 1500.0
			self.max_freq_offset = 50 * 1.25
			self.rrc_rolloff_rate, self.rrc_span = 0.9, 6
			pi_p, pi_gain = 0.4, 1800
		elif self.definition == '300':           # psk.py:26-55
			self.symbol_rate = 300.0
			self.input_bpf_low_cutoff, self.input_bpf_high_cutoff, self.input_bpf_span = 1200.0, 1800.0, 1.5
			self.carrier_freq = 1500.0
			self.max_freq_offset = 25 * 1.25
			self.rrc_rolloff_rate, self.rrc_span = 0.6, 6
			pi_p, pi_gain = 0.06, 7200
		else:
			raise ValueError(f"BPSKModem has no preset '{self.definition}' (psk.py:26, 56)")
		self.Loop_LPF = IIR_1(sample_rate=self.sample_rate, filter_type='lpf', cutoff=250.0, gain=1.0)
		self.FeedbackController = PI_control(p=pi_p, i=pi_p / 1000, i_limit=self.max_freq_offset, gain=pi_gain)
		self.oscillator_amplitude = 1.0
		self.tune()

	def retune(self, **kwargs):                  # psk.py:93-100
		for key in ('symbol_rate', 'input_bpf_low_cutoff', 'input_bpf_high_cutoff', 'input_bpf_span', 'sample_rate',
				'carrier_freq'):
			setattr(self, key, kwargs.get(key, getattr(self, key)))
		self.tune()

	def StringOptionsRetune(self, options):      # psk.py:102-109
		for key in ('symbol_rate', 'input_bpf_low_cutoff', 'input_bpf_high_cutoff', 'input_bpf_span', 'sample_rate',
				'carrier_freq'):
			setattr(self, key, float(options.get(key, getattr(self, key))))
		self.tune()

	def tune(self):                              # psk.py:111-160
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		self.AGC = AGC(sample_rate=self.sample_rate, attack_rate=self.agc_attack_rate, sustain_time=self.agc_sustain_time,
			decay_rate=self.agc_decay_rate, target_amplitude=self.oscillator_amplitude, record_envelope=False)
		self.NCO = NCO(sample_rate=self.sample_rate, amplitude=self.oscillator_amplitude,
			set_frequency=self.carrier_freq, wavetable_size=256)
		self.rrc = RRC(sample_rate=self.sample_rate, symbol_rate=self.symbol_rate, symbol_span=self.rrc_span,
			rolloff_rate=self.rrc_rolloff_rate)
		self.output_sample_rate = self.sample_rate

	def describe(self, desc, keep):
		loop = _lib.LoopDesc()
		keep.append(loop)
		self.AGC.describe(loop)
		self.NCO.describe(loop, keep)
		self.Loop_LPF.describe(loop)
		self.FeedbackController.describe(loop)
		desc.modem_kind = self.modem_kind
		desc.invert_soft = 0
		desc.bpf, desc.n_bpf = _arr(self.input_bpf, keep), len(self.input_bpf)
		desc.lpf, desc.n_lpf = _arr(self.rrc.taps, keep), len(self.rrc.taps)
		desc.loop = ctypes_pointer(loop)

	def demod(self, input_audio):
		"""psk.py:162-195 on the GPU -> float64 ndarray."""
		from ..engine import demod_only
		return demod_only(self, input_audio)


def ctypes_pointer(obj):
	import ctypes
	return ctypes.pointer(obj)


# config -> (constellation, agc attack, agc sustain, symbol_rate, bpf low, bpf high, bpf span [ms], hilbert span [ms],
#            carrier, max_freq_offset, rrc rolloff, loop cutoff, pi_p, pi_i divisor, pi gain)      psk.py:485-629
_MPSK_PRESETS = {
	'qpsk_3600': ('qpsk', 5000.0, 0.1, 1800, 300.0, 3000.0, 2, 4.5, 1650.0, 12.5 * 1.25, 0.3, 250.0, 0.15, 1000, (14400 / 65536)),
	'qpsk_600': ('qpsk', 500.0, 1, 300, 1200.0, 1800.0, 4, 3.4, 1500.0, 25, 0.6, 150, 0.1, 1000, (7200 / 65536)),
	'qpsk_2400': ('qpsk', 500.0, 1, 1200, 200.0, 2800.0, 2.7, 3.4, 1500.0, 25 * 1.25, 0.9, 250.0, 0.3, 2000, (14400 / 65536)),
	'bpsk_300': ('bpsk', 500.0, 1, 300, 1200.0, 1800.0, 2.7, 2.7, 1500.0, 50, 0.6, 250.0, 0.15, 1000, 1.5 * (500)),
	'bpsk_1200': ('bpsk', 500.0, 1, 1200, 200.0, 2800.0, 4.8, 2, 1500.0, 87.5, 0.9, 200.0, 0.15, 1000, 5),
}


class MPSKModem:
	modem_kind = _lib.PM_MODEM_MPSK

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', 'qpsk_3600')
		self.sample_rate = kwargs.get('sample_rate', 44100.0)
		if self.definition not in _MPSK_PRESETS:
			raise ValueError(f"MPSKModem has no preset '{self.definition}' (psk.py:485-629)")
		(self.constellation_id, self.agc_attack_rate, self.agc_sustain_time, self.symbol_rate, self.input_bpf_low_cutoff,
			self.input_bpf_high_cutoff, self.input_bpf_span, self.hilbert_span, self.carrier_freq, self.max_freq_offset,
			self.rrc_rolloff_rate, loop_cutoff, pi_p, pi_div, pi_gain) = _MPSK_PRESETS[self.definition]
		self.agc_decay_rate = 50.0
		self.rrc_span = 6
		self.Loop_LPF = IIR_1(sample_rate=self.sample_rate, filter_type='lpf', cutoff=loop_cutoff, gain=1)
		self.FeedbackController = PI_control(p=pi_p, i=pi_p / pi_div, i_limit=self.max_freq_offset, gain=pi_gain)
		self.oscillator_amplitude = 1.0
		self.pd_gain = 32
		self.tune()

	def StringOptionsRetune(self, options):      # psk.py:634-638
		self.symbol_rate = float(options.get('symbol_rate', self.symbol_rate))
		self.sample_rate = float(options.get('sample_rate', self.sample_rate))
		self.carrier_freq = float(options.get('carrier_freq', self.carrier_freq))
		self.tune()

	def tune(self):                              # psk.py:640-703
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / 1000)
		self.hilbert_tap_count = round(self.sample_rate * self.hilbert_span / 1000)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		if self.hilbert_tap_count % 2 == 0:
			self.hilbert_tap_count += 1
		self.Hilbert = Hilbert(tap_count=self.hilbert_tap_count)
		self.AGC = AGC(sample_rate=self.sample_rate, attack_rate=self.agc_attack_rate, sustain_time=self.agc_sustain_time,
			decay_rate=self.agc_decay_rate, target_amplitude=self.oscillator_amplitude, record_envelope=False)
		self.NCO = NCO(sample_rate=self.sample_rate, amplitude=self.oscillator_amplitude,
			set_frequency=self.carrier_freq, wavetable_size=256)
		self.rrc = RRC(sample_rate=self.sample_rate, symbol_rate=self.symbol_rate, symbol_span=self.rrc_span,
			rolloff_rate=self.rrc_rolloff_rate)
		self.output_sample_rate = self.sample_rate
		# psk.py:703: the loop starts at the maximum negative frequency offset
		self.FeedbackController.integral = -self.max_freq_offset

	def describe(self, desc, keep):
		loop = _lib.LoopDesc()
		keep.append(loop)
		self.AGC.describe(loop)
		self.NCO.describe(loop, keep)
		self.Loop_LPF.describe(loop)
		self.FeedbackController.describe(loop)
		PhaseDetector(self.constellation_id, 64, self.pd_gain).describe(loop, keep)      # psk.py:707
		self.Hilbert.describe(loop, keep)
		desc.modem_kind = self.modem_kind
		desc.invert_soft = 0
		desc.bpf, desc.n_bpf = _arr(self.input_bpf, keep), len(self.input_bpf)
		desc.lpf, desc.n_lpf = _arr(self.rrc.taps, keep), len(self.rrc.taps)
		desc.loop = ctypes_pointer(loop)

	def demod(self, input_audio):
		"""psk.py:705-773 on the GPU -> IQData of two float64 ndarrays."""
		from ..engine import demod_only
		return demod_only(self, input_audio)
