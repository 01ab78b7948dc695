"""CPU oracle for the pymodem demod_chain hot path -- TEST INFRASTRUCTURE ONLY.

A restatement of the reference's algorithm (NumPy for the FIR stages, the C
library built from oracle/oracle.c for the sequential / integer stages).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product path (pymodem_b200/) never does.

Parity pin: the reference ships no tests; this oracle is
pinned against the live reference imported from /root/reference in the build
container (tools/make_golden.py -> tests/golden/, tests/test_oracle_golden.py).

Reference citations are paths under /root/reference/modems_codecs/.
"""
import ctypes
import os
import subprocess
from math import ceil

import numpy as np
from scipy.signal import firwin

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
	"""Compile oracle/oracle.c with gcc (oracle/Makefile)."""
	src = os.path.join(_HERE, "oracle.c")
	if (not force and os.path.exists(_LIB_PATH)
			and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
		return _LIB_PATH
	subprocess.run(["make", "-C", _HERE, "-B", "_build/liboracle.so"], check=True,
		stdout=subprocess.DEVNULL)
	return _LIB_PATH


class _Slicer(ctypes.Structure):
	_fields_ = [
		("phase_clock", ctypes.c_double),
		("samples_per_symbol", ctypes.c_double),
		("rollover_threshold", ctypes.c_double),
		("lock_rate", ctypes.c_double),
		("last_sample", ctypes.c_double),
		("last_q_sample", ctypes.c_double),
		("streamaddress", ctypes.c_int64),
		("working_byte", ctypes.c_uint32),
		("working_bit_count", ctypes.c_uint32),
		("state_register", ctypes.c_uint32),
		("state_mask", ctypes.c_uint32),
		("bits_per_symbol", ctypes.c_uint32),
		("demap", ctypes.c_uint32 * 16),
	]


class _Lfsr(ctypes.Structure):
	_fields_ = [
		("polynomial", ctypes.c_uint64),
		("shift_register", ctypes.c_uint64),
		("invert", ctypes.c_int),
	]


class _Loop(ctypes.Structure):
	"""orc_loop"""
	_fields_ = [(n, ctypes.c_double) for n in (
		"agc_scaled_attack", "agc_scaled_decay", "agc_sustain_time", "agc_sustain_increment", "agc_target",
		"nco_phase_scale", "nco_index_scale", "nco_set_frequency", "nco_two_pi", "nco_quarter")] + [
		("nco_wavetable", ctypes.c_void_p), ("nco_size", ctypes.c_int64),
		("iir_b0", ctypes.c_double), ("iir_b1", ctypes.c_double), ("iir_a1", ctypes.c_double),
		("pi_gain", ctypes.c_double), ("pi_p", ctypes.c_double), ("pi_i", ctypes.c_double),
		("pi_limit", ctypes.c_double), ("pi_integral0", ctypes.c_double),
		("pd_table", ctypes.c_void_p), ("pd_granularity", ctypes.c_int64)]


def lib():
	global _lib
	if _lib is None:
		build()
		L = ctypes.CDLL(_LIB_PATH)
		P = ctypes.POINTER
		L.orc_slicer_init.argtypes = [P(_Slicer), ctypes.c_double, ctypes.c_double, ctypes.c_double]
		L.orc_qslicer_init.argtypes = [P(_Slicer), ctypes.c_double, ctypes.c_double, ctypes.c_double,
			ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
		L.orc_binary_slice.restype = ctypes.c_int64
		L.orc_binary_slice.argtypes = [P(_Slicer), ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
		L.orc_quadrature_slice.restype = ctypes.c_int64
		L.orc_quadrature_slice.argtypes = [P(_Slicer), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
		L.orc_lfsr_init.argtypes = [P(_Lfsr), ctypes.c_uint64, ctypes.c_int]
		L.orc_lfsr_unscramble.argtypes = [P(_Lfsr), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
		L.orc_check_crc.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
		L.orc_crc16.restype = ctypes.c_uint32
		L.orc_crc16.argtypes = [ctypes.c_void_p, ctypes.c_int64]
		L.orc_validate_header.restype = ctypes.c_int
		L.orc_validate_header.argtypes = [ctypes.c_void_p, ctypes.c_int64]
		L.orc_ax25_new.restype = ctypes.c_void_p
		L.orc_ax25_free.argtypes = [ctypes.c_void_p]
		L.orc_ax25_decode.restype = ctypes.c_int64
		L.orc_ax25_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
		L.orc_gf_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
		L.orc_rs_genpoly.restype = ctypes.c_int
		L.orc_rs_genpoly.argtypes = [ctypes.c_int, ctypes.c_void_p]
		L.orc_rs_decode.restype = ctypes.c_int
		L.orc_rs_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
		L.orc_il2p_new.restype = ctypes.c_void_p
		L.orc_il2p_new.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
		L.orc_il2p_free.argtypes = [ctypes.c_void_p]
		L.orc_il2p_decode.restype = ctypes.c_int64
		L.orc_il2p_decode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
		L.orc_buffer_max.restype = ctypes.c_double
		L.orc_buffer_max.argtypes = [ctypes.c_void_p, ctypes.c_int64]
		L.orc_agc_apply.argtypes = [P(_Loop), ctypes.c_void_p, ctypes.c_int64]
		L.orc_bpsk_loop.argtypes = [P(_Loop), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
		L.orc_pll_loop.argtypes = [P(_Loop), ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
		L.orc_mpsk_loop.argtypes = [P(_Loop), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
			ctypes.c_void_p, ctypes.c_void_p]
		_lib = L
	return _lib


def check_boolean(s):
	"""string_ops.py:6-15"""
	return s.lower() in ("yes", "true", "1")


def _ptr(a):
	return a.ctypes.data_as(ctypes.c_void_p)


# ----------------------------------------------------------------------------
# Modems (FIR stages in NumPy, the same numpy.convolve calls the reference makes)
# ----------------------------------------------------------------------------
class AFSKModem:
	"""afsk.py:13-167"""

	def __init__(self, sample_rate, config, options):
		self.sample_rate = sample_rate
		if config == '300':                         # afsk.py:19-42
			p = dict(symbol_rate=300.0, input_bpf_low_cutoff=1500.0, input_bpf_high_cutoff=1900.0,
				input_bpf_span=7, mark_freq=1695.0, space_freq=1705.0, space_gain=1.0,
				output_lpf_cutoff=240.0, output_lpf_span=2.5, correlator_span=0.3, correlator_offset=0.0)
		else:                                       # afsk.py:43-66
			p = dict(symbol_rate=1200.0, input_bpf_low_cutoff=900.0, input_bpf_high_cutoff=2500.0,
				input_bpf_span=3.7, mark_freq=1200.0, space_freq=2200.0, space_gain=1.0,
				output_lpf_cutoff=1400.0, output_lpf_span=2.5, correlator_span=1.0, correlator_offset=0.0)
		p['sample_rate'] = sample_rate
		for key in list(p):                         # afsk.py:87-99 (options are strings)
			p[key] = float(options.get(key, p[key]))
		self.__dict__.update(p)
		self.tune()

	def tune(self):                                 # afsk.py:102-146
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / self.symbol_rate)
		self.output_lpf_tap_count = round(self.sample_rate * self.output_lpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count,
			[self.input_bpf_low_cutoff, self.input_bpf_high_cutoff], pass_zero='bandpass', fs=self.sample_rate)
		self.output_lpf = firwin(self.output_lpf_tap_count, self.output_lpf_cutoff, fs=self.sample_rate)
		time_indices = np.arange(ceil(self.correlator_span * self.sample_rate / self.symbol_rate))
		mark_indices = time_indices * (2.0 * np.pi * (self.mark_freq + self.correlator_offset) / self.sample_rate)
		self.mark_correlator_i = np.cos(mark_indices)
		self.mark_correlator_q = np.sin(mark_indices)
		space_indices = time_indices * (2.0 * np.pi * (self.space_freq + self.correlator_offset) / self.sample_rate)
		self.space_correlator_i = self.space_gain * np.cos(space_indices)
		self.space_correlator_q = self.space_gain * np.sin(space_indices)
		self.output_sample_rate = 1.0 * self.sample_rate

	@property
	def halo(self):
		return (self.input_bpf_tap_count - 1) + (len(self.mark_correlator_i) - 1) + (self.output_lpf_tap_count - 1)

	def demod(self, input_audio):                   # afsk.py:148-167
		audio = np.convolve(input_audio, self.input_bpf, 'valid')
		mark_mag = np.sqrt(np.convolve(audio, self.mark_correlator_i, 'valid')**2
			+ np.convolve(audio, self.mark_correlator_q, 'valid')**2)
		space_mag = np.sqrt(np.convolve(audio, self.space_correlator_i, 'valid')**2
			+ np.convolve(audio, self.space_correlator_q, 'valid')**2)
		audio = mark_mag - space_mag
		return np.convolve(audio, self.output_lpf, 'valid')


class FSKModem:
	"""fsk.py:15-159 ('lpf' presets; the rrc presets are built from rrc.py)"""

	def __init__(self, sample_rate, config, options):
		self.sample_rate = sample_rate
		self.rrc_rolloff_rate = False
		self.invert = False
		if config == '4800':                        # fsk.py:36-44
			self.symbol_rate, self.input_filter_type = 4800.0, 'lpf'
			self.input_lpf_cutoff, self.input_lpf_span = 3000.0, 1.5
		elif config == '4800-rrc':
			self.symbol_rate, self.input_filter_type = 4800.0, 'rrc'
			self.rrc_rolloff_rate, self.input_lpf_span = 0.2, 9
		elif config == '9600-rrc':
			self.symbol_rate, self.input_filter_type = 9600.0, 'rrc'
			self.rrc_rolloff_rate, self.input_lpf_span = 0.2, 9
		elif config == '4800-gauss':
			self.symbol_rate, self.input_filter_type = 4800.0, 'lpf'
			self.input_lpf_cutoff, self.input_lpf_span = 0.9 * 4800.0, 4
		elif config == '9600-gauss':
			self.symbol_rate, self.input_filter_type = 9600.0, 'lpf'
			self.input_lpf_cutoff, self.input_lpf_span = 0.9 * 9600.0, 4
		else:                                       # '9600' and default, fsk.py:25-35, 92-103
			self.symbol_rate, self.input_filter_type = 9600.0, 'lpf'
			self.input_lpf_cutoff, self.input_lpf_span = 6000.0, 1.5
		self.invert = check_boolean(options.get('invert', "false"))   # fsk.py:110-113
		self.tune()

	def tune(self):                                 # fsk.py:115-147
		self.input_lpf_tap_count = round(self.sample_rate * self.input_lpf_span / self.symbol_rate)
		if self.input_filter_type == 'rrc':
			self.input_lpf = rrc_taps(self.sample_rate, self.symbol_rate, self.input_lpf_span, self.rrc_rolloff_rate)
		else:
			self.input_lpf = firwin(self.input_lpf_tap_count, [self.input_lpf_cutoff],
				pass_zero='lowpass', fs=self.sample_rate)

	@property
	def halo(self):
		return len(self.input_lpf) - 1

	def demod(self, input_audio):                   # fsk.py:149-159
		audio = np.convolve(input_audio, self.input_lpf, 'valid')
		if self.invert:
			audio = -audio
		return audio


def rrc_taps(sample_rate, symbol_rate, symbol_span, rolloff_rate):
	"""rrc.py:18-95 with the default 'rect' window (the only one ever selected)."""
	import math
	oversample = sample_rate / symbol_rate
	tap_count = int(round(symbol_span * oversample, 0)) + 1
	time_step = 1 / sample_rate
	symbol_time = 1 / symbol_rate
	time = np.arange(0, tap_count * time_step, time_step) - (tap_count * time_step / 2) + (time_step / 2)
	taps = []
	try:
		asymptote = symbol_time / (4 * rolloff_rate)
	except ZeroDivisionError:
		asymptote = False
	for t in time:
		if math.isclose(t, -asymptote) or math.isclose(t, asymptote):
			numerator = rolloff_rate * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * rolloff_rate))
				+ (1 - (2 / np.pi)) * np.cos(np.pi / (4 * rolloff_rate)))
			denominator = symbol_time * pow(2, 0.5)
			taps.append(numerator / denominator)
		else:
			numerator = np.sin(np.pi * t * (1 - rolloff_rate) / symbol_time) \
				+ 4 * rolloff_rate * t * np.cos(np.pi * t * (1 + rolloff_rate) / symbol_time) / symbol_time
			denominator = np.pi * t * (1 - pow(4 * rolloff_rate * t / symbol_time, 2)) / symbol_time
			taps.append(numerator / (denominator * symbol_time))
	taps = taps / np.linalg.norm(taps)
	return np.multiply(taps, [1] * len(taps))


# ----------------------------------------------------------------------------
# PSK / PLL modems: FIRs in NumPy, AGC and the carrier loops in C
# ----------------------------------------------------------------------------
def _loop_desc(sample_rate, agc, carrier_freq, iir_cutoff, iir_gain, pi_p, pi_i, pi_limit, pi_gain,
		integral0=0.0, with_pd=False, keep=None):
	"""Constants of agc.py:8-24, nco.py:14-32, iir.py:15-29, pi_control.py:8-13, phase_detector.py:12-45
	computed with the reference's own expressions (math.sin / math.tan / atan2 of the host libm)."""
	from math import atan2, pi, sin, sqrt, tan
	L = _Loop()
	attack_rate, sustain_time, decay_rate, target = agc
	L.agc_scaled_attack = attack_rate / sample_rate
	L.agc_scaled_decay = decay_rate / sample_rate
	L.agc_sustain_time = sustain_time
	L.agc_sustain_increment = 1 / sample_rate
	L.agc_target = target
	size = 256
	amplitude = 1.0
	wavetable = np.array([amplitude * sin(i * 2.0 * pi / size) for i in range(size)], dtype=np.float64)
	keep.append(wavetable)
	L.nco_wavetable = wavetable.ctypes.data
	L.nco_size = size
	L.nco_index_scale = size / (2.0 * pi)
	L.nco_phase_scale = 2.0 * pi / sample_rate
	L.nco_set_frequency = carrier_freq
	L.nco_two_pi = 2.0 * pi
	L.nco_quarter = size / 4.0
	radian_cutoff = 2.0 * pi * iir_cutoff                       # iir.py:15-26
	warp_cutoff = 2.0 * sample_rate * tan(radian_cutoff / (2.0 * sample_rate))
	omega_T = warp_cutoff / sample_rate
	a1 = (2.0 - omega_T) / (2.0 + omega_T)
	b0 = omega_T / (2.0 + omega_T)
	L.iir_b0, L.iir_b1, L.iir_a1 = iir_gain * b0, iir_gain * b0, a1
	L.pi_gain, L.pi_p, L.pi_i, L.pi_limit, L.pi_integral0 = pi_gain, pi_p, pi_i, pi_limit, integral0
	if with_pd:                                                 # phase_detector.py:12-45 ('qpsk', 64, 32)
		g, gain = 64, 32
		min_mag, max_mag = g * .15, g * .76
		table = np.zeros((g, g), dtype=np.int32)
		for real in range(g):
			for imag in range(g):
				mag = sqrt((real**2) + (imag**2))
				if mag >= min_mag and mag <= max_mag:
					table[real, imag] = round(gain * ((atan2(imag, real) * 180 / pi) - 45))
		keep.append(table)
		L.pd_table = table.ctypes.data
		L.pd_granularity = g
	return L


def hilbert_taps(tap_count):
	"""hilbert.py:9-34 (hann window) -> (taps, delay)"""
	from math import pi, sin
	delay = tap_count // 2
	taps = []
	for n in range(-delay, -delay + tap_count):
		taps.append(2 / (pi * n) if n % 2 else 0)
	N = tap_count - 1
	return np.array([taps[i] * sin(pi * i / N)**2 for i in range(tap_count)], dtype=np.float64), delay


class BPSKModem:
	"""psk.py:20-195"""

	def __init__(self, sample_rate, config, options):
		if config == '1200':                        # psk.py:56-88
			p = dict(symbol_rate=1200.0, input_bpf_low_cutoff=200.0, input_bpf_high_cutoff=2800.0, input_bpf_span=4.80,
				carrier_freq=1500.0)
			self.max_freq_offset, self.rrc_rolloff_rate, self.rrc_span = 50 * 1.25, 0.9, 6
			self.iir_cutoff, pi_p, self.pi_gain = 250.0, 0.4, 1800
		else:                                       # '300', psk.py:26-55
			p = dict(symbol_rate=300.0, input_bpf_low_cutoff=1200.0, input_bpf_high_cutoff=1800.0, input_bpf_span=1.5,
				carrier_freq=1500.0)
			self.max_freq_offset, self.rrc_rolloff_rate, self.rrc_span = 25 * 1.25, 0.6, 6
			self.iir_cutoff, pi_p, self.pi_gain = 250.0, 0.06, 7200
		self.pi_p, self.pi_i = pi_p, pi_p / 1000
		p['sample_rate'] = sample_rate
		for key in list(p):                         # psk.py:102-109
			p[key] = float(options.get(key, p[key]))
		self.__dict__.update(p)
		self.tune()

	def tune(self):                                 # psk.py:111-160
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		self.rrc = rrc_taps(self.sample_rate, self.symbol_rate, self.rrc_span, self.rrc_rolloff_rate)
		self._keep = []
		self.loop = _loop_desc(self.sample_rate, (500.0, 1.0, 50.0, 1.0), self.carrier_freq, self.iir_cutoff, 1.0,
			self.pi_p, self.pi_i, self.max_freq_offset, self.pi_gain, keep=self._keep)
		self.output_sample_rate = self.sample_rate

	def demod(self, input_audio):                   # psk.py:162-195
		audio = np.ascontiguousarray(np.convolve(input_audio, self.input_bpf, 'valid'), dtype=np.float64)
		lib().orc_agc_apply(ctypes.byref(self.loop), _ptr(audio), len(audio))
		demod_audio = np.empty_like(audio)
		lib().orc_bpsk_loop(ctypes.byref(self.loop), _ptr(audio), len(audio), _ptr(demod_audio))
		self.loop_input = audio
		self.loop_out = demod_audio
		return np.convolve(demod_audio, self.rrc, 'valid')


class AFSKPLLModem:
	"""afsk_pll.py:16-170"""

	def __init__(self, sample_rate, config, options):
		p = dict(symbol_rate=300.0, input_bpf_low_cutoff=1500.0, input_bpf_high_cutoff=1900.0, input_bpf_span=7.0,
			output_lpf_cutoff=240.0, output_lpf_span=5, carrier_freq=1700.0)     # afsk_pll.py:22-50
		p['sample_rate'] = sample_rate
		for key in list(p):                         # afsk_pll.py:70-79
			p[key] = float(options.get(key, p[key]))
		self.__dict__.update(p)
		self.tune()

	def tune(self):                                 # afsk_pll.py:81-138
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / self.symbol_rate)
		self.output_lpf_tap_count = round(self.sample_rate * self.output_lpf_span / self.symbol_rate)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		self.output_lpf = firwin(self.output_lpf_tap_count, self.output_lpf_cutoff, fs=self.sample_rate, scale=True)
		self._keep = []
		pi_p = 0.6
		self.loop = _loop_desc(self.sample_rate, (500.0, 1.0, 50.0, 1.0), self.carrier_freq, 150.0, 1.0,
			pi_p, pi_p / 6000, 50, 900, keep=self._keep)
		self.output_sample_rate = self.sample_rate

	def demod(self, input_audio):                   # afsk_pll.py:140-170
		audio = np.ascontiguousarray(np.convolve(input_audio, self.input_bpf, 'valid'), dtype=np.float64)
		lib().orc_agc_apply(ctypes.byref(self.loop), _ptr(audio), len(audio))
		demod_audio = np.empty_like(audio)
		lib().orc_pll_loop(ctypes.byref(self.loop), _ptr(audio), len(audio), _ptr(demod_audio))
		self.loop_input = audio
		self.loop_out = demod_audio
		return np.convolve(demod_audio, self.output_lpf, 'valid')


class MPSKModem:
	"""psk.py:479-773; demod() returns (i_data, q_data)"""
	# config -> (agc attack, sustain, symbol_rate, bpf low, high, bpf span ms, hilbert span ms, carrier,
	#            max_freq_offset, rrc rolloff, iir cutoff, pi_p, pi_i divisor, pi gain)   psk.py:485-629
	_PRESETS = {
		'qpsk_3600': (5000.0, 0.1, 1800, 300.0, 3000.0, 2, 4.5, 1650.0, 12.5 * 1.25, 0.3, 250.0, 0.15, 1000, (14400 / 65536)),
		'qpsk_600': (500.0, 1, 300, 1200.0, 1800.0, 4, 3.4, 1500.0, 25, 0.6, 150, 0.1, 1000, (7200 / 65536)),
		'qpsk_2400': (500.0, 1, 1200, 200.0, 2800.0, 2.7, 3.4, 1500.0, 25 * 1.25, 0.9, 250.0, 0.3, 2000, (14400 / 65536)),
		'bpsk_300': (500.0, 1, 300, 1200.0, 1800.0, 2.7, 2.7, 1500.0, 50, 0.6, 250.0, 0.15, 1000, 1.5 * (500)),
		'bpsk_1200': (500.0, 1, 1200, 200.0, 2800.0, 4.8, 2, 1500.0, 87.5, 0.9, 200.0, 0.15, 1000, 5),
	}

	def __init__(self, sample_rate, config, options):
		(self.agc_attack_rate, self.agc_sustain_time, symbol_rate, self.input_bpf_low_cutoff, self.input_bpf_high_cutoff,
			self.input_bpf_span, self.hilbert_span, carrier_freq, self.max_freq_offset, self.rrc_rolloff_rate,
			self.iir_cutoff, self.pi_p, pi_div, self.pi_gain) = self._PRESETS[config]
		self.pi_i = self.pi_p / pi_div
		self.rrc_span = 6
		# psk.py:634-638
		self.symbol_rate = float(options.get('symbol_rate', symbol_rate))
		self.sample_rate = float(options.get('sample_rate', sample_rate))
		self.carrier_freq = float(options.get('carrier_freq', carrier_freq))
		self.tune()

	def tune(self):                                 # psk.py:640-703
		self.input_bpf_tap_count = round(self.sample_rate * self.input_bpf_span / 1000)
		self.hilbert_tap_count = round(self.sample_rate * self.hilbert_span / 1000)
		self.input_bpf = firwin(self.input_bpf_tap_count, [self.input_bpf_low_cutoff, self.input_bpf_high_cutoff],
			pass_zero='bandpass', fs=self.sample_rate, scale=True)
		if self.hilbert_tap_count % 2 == 0:
			self.hilbert_tap_count += 1
		self.hilbert, self.hilbert_delay = hilbert_taps(self.hilbert_tap_count)
		self.rrc = rrc_taps(self.sample_rate, self.symbol_rate, self.rrc_span, self.rrc_rolloff_rate)
		self._keep = []
		self.loop = _loop_desc(self.sample_rate, (self.agc_attack_rate, self.agc_sustain_time, 50.0, 1.0), self.carrier_freq,
			self.iir_cutoff, 1.0, self.pi_p, self.pi_i, self.max_freq_offset, self.pi_gain,
			integral0=-self.max_freq_offset, with_pd=True, keep=self._keep)      # psk.py:703
		self.output_sample_rate = self.sample_rate

	def demod(self, input_audio):                   # psk.py:705-773
		audio = np.ascontiguousarray(np.convolve(input_audio, self.input_bpf, 'valid'), dtype=np.float64)
		lib().orc_agc_apply(ctypes.byref(self.loop), _ptr(audio), len(audio))
		imag_audio = np.ascontiguousarray(np.convolve(audio, self.hilbert, 'valid'))
		delay_taps = [0] * (self.hilbert_delay + 1)
		delay_taps[0] = 1
		real_audio = np.convolve(audio, delay_taps, 'valid')
		real_audio = np.ascontiguousarray(real_audio[:-self.hilbert_delay])
		n = min(len(real_audio), len(imag_audio))
		i_data = np.empty(n, dtype=np.float64)
		q_data = np.empty(n, dtype=np.float64)
		lib().orc_mpsk_loop(ctypes.byref(self.loop), _ptr(real_audio), _ptr(imag_audio), n, _ptr(i_data), _ptr(q_data))
		self.loop_out = (i_data, q_data)
		return np.convolve(i_data, self.rrc, 'valid'), np.convolve(q_data, self.rrc, 'valid')


_QPSK_DEMAP = [3, 1, 2, 0, 2, 3, 0, 1, 1, 0, 3, 2, 0, 2, 1, 3]
_BPSK_DEMAP = [0, 0, 1, 1]


class QuadratureSlicer:
	"""slicer.py:109-242"""
	_PRESETS = {                                    # slicer.py:124-165
		'qpsk_600': (0xF, 2, _QPSK_DEMAP, 300, 0.815), 'bpsk_300': (0x3, 1, _BPSK_DEMAP, 300, 0.815),
		'bpsk_1200': (0x3, 1, _BPSK_DEMAP, 1200, 0.9), 'qpsk_2400': (0xF, 2, _QPSK_DEMAP, 1200, 0.9),
		'qpsk_4800': (0xF, 2, _QPSK_DEMAP, 2400, 0.99), 'qpsk_3600': (0xF, 2, _QPSK_DEMAP, 1800, 0.99),
	}

	def __init__(self, sample_rate, config, options):
		mask, bps, demap, symbol_rate, lock_rate = self._PRESETS.get(config, (0xF, 2, _QPSK_DEMAP, 1200, 0.9))
		lock_rate = float(options.get('lock_rate', lock_rate))      # slicer.py:176-180
		self.state = _Slicer()
		d = (ctypes.c_uint32 * 16)(*demap)
		lib().orc_qslicer_init(ctypes.byref(self.state), float(sample_rate), float(symbol_rate), lock_rate, mask, bps, d)
		self.bits_per_symbol = bps

	def slice(self, iq):
		i_s = np.ascontiguousarray(iq[0], dtype=np.float64)
		q_s = np.ascontiguousarray(iq[1], dtype=np.float64)
		n = min(len(i_s), len(q_s))
		cap = int(n / max(self.state.rollover_threshold, 1.0) * self.bits_per_symbol / 8) + 16
		out_b = np.empty(cap, dtype=np.uint8)
		out_a = np.empty(cap, dtype=np.int64)
		cnt = lib().orc_quadrature_slice(ctypes.byref(self.state), _ptr(i_s), _ptr(q_s), n, _ptr(out_b), _ptr(out_a), cap)
		assert cnt <= cap
		return out_b[:cnt], out_a[:cnt]


# ----------------------------------------------------------------------------
# Sequential / integer stages (C)
# ----------------------------------------------------------------------------
class BinarySlicer:
	"""slicer.py:9-107"""

	def __init__(self, sample_rate, config, options):
		if config == '300':
			symbol_rate, lock_rate = 300, 0.75
		elif config == '9600':
			symbol_rate, lock_rate = 9600, 0.88
		elif config == '4800':
			symbol_rate, lock_rate = 4800, 0.88
		else:
			symbol_rate, lock_rate = 1200, 0.75
		lock_rate = float(options.get('lock_rate', lock_rate))      # slicer.py:46
		self.symbol_rate, self.lock_rate, self.sample_rate = symbol_rate, lock_rate, sample_rate
		self.state = _Slicer()
		lib().orc_slicer_init(ctypes.byref(self.state), float(sample_rate), float(symbol_rate), lock_rate)

	def slice(self, samples):
		samples = np.ascontiguousarray(samples, dtype=np.float64)
		n = len(samples)
		cap = int(n / max(self.state.rollover_threshold, 1.0) / 8) + 16
		out_b = np.empty(cap, dtype=np.uint8)
		out_a = np.empty(cap, dtype=np.int64)
		cnt = lib().orc_binary_slice(ctypes.byref(self.state), _ptr(samples), n, _ptr(out_b), _ptr(out_a), cap)
		assert cnt <= cap
		return out_b[:cnt], out_a[:cnt]


class LFSR:
	"""lfsr.py:10-52"""

	def __init__(self, options):
		self.polynomial = int(options.get('poly', '0x1'), 16)       # lfsr.py:19
		self.invert = check_boolean(options.get('invert', "false"))
		self.state = _Lfsr()
		lib().orc_lfsr_init(ctypes.byref(self.state), self.polynomial, int(self.invert))

	def stream_unscramble_8bit(self, data, addr):
		data = np.ascontiguousarray(data, dtype=np.uint8)
		out = np.empty_like(data)
		lib().orc_lfsr_unscramble(ctypes.byref(self.state), _ptr(data), _ptr(out), len(data))
		return out, addr


class AX25Codec:
	"""ax25.py:11-93.  decode() returns [(streamaddress, bytes, BytesCorrected=0), ...]"""

	def __init__(self, ident):
		self.identifier = ident
		self._h = lib().orc_ax25_new()

	def __del__(self):
		try:
			lib().orc_ax25_free(self._h)
		except Exception:
			pass

	def decode(self, data, addr):
		data = np.ascontiguousarray(data, dtype=np.uint8)
		addr = np.ascontiguousarray(addr, dtype=np.int64)
		n = len(data)
		rec_cap, arena_cap = 1024, 1 << 16
		while True:
			# decode mutates codec state, so size buffers pessimistically: at
			# most one packet per 19 stream bytes, data never longer than the
			# bytes seen so far plus what was pending.
			rec_cap = max(rec_cap, n // 19 + 4)
			arena_cap = max(arena_cap, 2 * n + (1 << 16))
			rec_addr = np.empty(rec_cap, dtype=np.int64)
			rec_off = np.empty(rec_cap, dtype=np.int64)
			rec_len = np.empty(rec_cap, dtype=np.int64)
			arena = np.empty(arena_cap, dtype=np.uint8)
			used = ctypes.c_int64(0)
			nrec = lib().orc_ax25_decode(self._h, _ptr(data), _ptr(addr), n, _ptr(rec_addr), _ptr(rec_off),
				_ptr(rec_len), rec_cap, _ptr(arena), arena_cap, ctypes.byref(used))
			if nrec <= rec_cap and used.value <= arena_cap:
				break
			raise RuntimeError("oracle AX.25 buffers too small (pending data longer than input)")
		return [(int(rec_addr[r]), bytes(arena[rec_off[r]:rec_off[r] + rec_len[r]]), 0) for r in range(nrec)]


class IL2PCodec:
	"""il2p.py:109-519.  decode() returns [(streamaddress, bytes, BytesCorrected), ...]"""

	def __init__(self, ident, options=None):
		options = options or {}
		self.identifier = ident
		# il2p.py:140-145 StringOptionsRetune (chain_builder.py:66 always calls it)
		self.collect_trailing_crc = check_boolean(options.get('crc', 'yes'))
		self.disable_rs = check_boolean(options.get('disable_rs', 'no'))
		self.min_distance = int(options.get('min_dist', 0))
		self.sync_tolerance = int(options.get('sync_tol', 0))
		self._h = lib().orc_il2p_new(int(self.collect_trailing_crc), int(self.disable_rs), self.min_distance,
			self.sync_tolerance)

	def __del__(self):
		try:
			lib().orc_il2p_free(self._h)
		except Exception:
			pass

	def decode(self, data, addr):
		data = np.ascontiguousarray(data, dtype=np.uint8)
		addr = np.ascontiguousarray(addr, dtype=np.int64)
		n = len(data)
		rec_cap = n // 17 + 16            # shortest frame: 3 sync bytes + a 15-byte header (no payload, no trailing CRC)
		arena_cap = 2 * n + (1 << 16)
		rec_addr = np.empty(rec_cap, dtype=np.int64)
		rec_off = np.empty(rec_cap, dtype=np.int64)
		rec_len = np.empty(rec_cap, dtype=np.int64)
		rec_corr = np.empty(rec_cap, dtype=np.int64)
		arena = np.empty(arena_cap, dtype=np.uint8)
		used = ctypes.c_int64(0)
		nrec = lib().orc_il2p_decode(self._h, _ptr(data), _ptr(addr), n, _ptr(rec_addr), _ptr(rec_off), _ptr(rec_len),
			_ptr(rec_corr), rec_cap, _ptr(arena), arena_cap, ctypes.byref(used))
		if nrec > rec_cap or used.value > arena_cap:
			raise RuntimeError("oracle IL2P buffers too small")
		return [(int(rec_addr[r]), bytes(arena[rec_off[r]:rec_off[r] + rec_len[r]]), int(rec_corr[r])) for r in range(nrec)]


def gf_tables():
	"""gf_functions.initialize(8, 0x11D) -> (table[255], index[256], inverse[256])"""
	t, i, v = np.zeros(255, dtype=np.int32), np.zeros(256, dtype=np.int32), np.zeros(256, dtype=np.int32)
	lib().orc_gf_tables(_ptr(t), _ptr(i), _ptr(v))
	return t, i, v


def rs_genpoly(num_roots):
	"""rs_functions.initialize(0, num_roots, 8, 0x11D)['genpoly']"""
	out = np.zeros(num_roots + 1, dtype=np.int32)
	lib().orc_rs_genpoly(num_roots, _ptr(out))
	return [int(x) for x in out]


def rs_decode(num_roots, data, min_distance=0):
	"""rs_functions.decode -> (result, corrected data)"""
	buf = np.array(list(data), dtype=np.uint8)
	r = lib().orc_rs_decode(num_roots, _ptr(buf), len(buf), min_distance)
	return int(r), bytes(buf)


def check_crc(data):
	"""crc_functions.py:9-61 -> [carried, calculated, valid]"""
	a = np.frombuffer(bytes(data), dtype=np.uint8)
	out = np.zeros(3, dtype=np.uint32)
	lib().orc_check_crc(_ptr(a), len(a), _ptr(out))
	return [int(out[0]), int(out[1]), bool(out[2])]


def crc16(data):
	a = np.frombuffer(bytes(data), dtype=np.uint8)
	return int(lib().orc_crc16(_ptr(a), len(a)))


def validate_header(data):
	"""packet_meta.py:21-41"""
	a = np.frombuffer(bytes(data), dtype=np.uint8)
	return bool(lib().orc_validate_header(_ptr(a), len(a)))


# ----------------------------------------------------------------------------
# Chain assembly (chain_builder.py:17-69, pymodem.py:68-114) and execution
# (chain_execute.py:6-28)
# ----------------------------------------------------------------------------
class Chain:
	def __init__(self, sample_rate, line):
		self.name = line['object_name']
		m = line['modem']
		if m['type'] == 'afsk':
			self.modem = AFSKModem(sample_rate, m['config'], m['options'])
		elif m['type'] == 'fsk':
			self.modem = FSKModem(sample_rate, m['config'], m['options'])
		elif m['type'] == 'bpsk':
			self.modem = BPSKModem(sample_rate, m['config'], m['options'])
		elif m['type'] == 'mpsk':
			self.modem = MPSKModem(sample_rate, m['config'], m['options'])
		elif m['type'] == 'afsk_pll':
			self.modem = AFSKPLLModem(sample_rate, m['config'], m['options'])
		else:
			raise NotImplementedError(f"oracle modem type {m['type']}")
		# pymodem.py:86-90: slicer runs at modem.output_sample_rate if it has one
		slicer_rate = getattr(self.modem, 'output_sample_rate', sample_rate)
		s = line['slicer']
		if s['type'] == 'binary':
			self.slicer = BinarySlicer(slicer_rate, s['config'], s['options'])
		elif s['type'] == 'quadrature':
			self.slicer = QuadratureSlicer(slicer_rate, s['config'], s['options'])
		else:
			raise NotImplementedError(f"oracle slicer type {s['type']}")
		self.stream = LFSR(line['stream']['options'])
		c = line['codec']
		if c['type'].lower() == 'ax25':
			self.codec = AX25Codec(self.name)
		elif c['type'].lower() == 'il2p':
			self.codec = IL2PCodec(self.name, c.get('options', {}))      # chain_builder.py:64-66
		else:
			raise NotImplementedError(f"oracle codec type {c['type']}")

	def process(self, audio):
		"""One call = chain_execute.process_chain; blocks keep state across calls."""
		soft = self.modem.demod(audio)
		b, a = self.slicer.slice(soft)
		b, a = self.stream.stream_unscramble_8bit(b, a)
		return self.codec.decode(b, a)

	def process_chunked(self, audio, chunk=1 << 20):
		"""Streaming form for long audio: FIRs are 'valid', so
		feeding audio[pos:pos+chunk+halo] and advancing by chunk is bit-identical
		to the monolithic call while slicer/LFSR/codec carry their state."""
		out = []
		halo = self.modem.halo
		n = len(audio)
		pos = 0
		while pos + halo < n:
			out += self.process(audio[pos:min(n, pos + chunk + halo)])
			pos += chunk
		return out


def build_chains(sample_rate, config_lines):
	return [Chain(sample_rate, line) for line in config_lines if line.get('object_type') == 'demod_chain']


def run_config(sample_rate, config_lines, audio, chunk=None):
	"""Deterministic driver: chains in config order. -> list (per chain) of
	[(streamaddress, data bytes, BytesCorrected), ...]"""
	chains = build_chains(sample_rate, config_lines)
	if chunk:
		return [c.process_chunked(audio, chunk) for c in chains]
	return [c.process(audio) for c in chains]


def correlate(per_chain, names, address_distance):
	"""PacketMetaArray.CalcCRCs + Correlate, packet_meta.py:219-271.
	-> (unique list of (streamaddress, data, crc, [decoders]), bad_count)"""
	uniques = []
	bad = 0
	first = True
	for name, packets in zip(names, per_chain):
		for (addr, data, _bc) in packets:
			carried, calc, valid = check_crc(data)
			if not (valid and validate_header(data)):
				bad += 1
				continue
			is_unique = True
			if not first:
				for u in uniques:
					if u[4] != name and abs(addr - u[0]) < address_distance and calc == u[2]:
						is_unique = False
						u[3].append(name)
						break
			if is_unique:
				uniques.append([addr, data, calc, [name], name])
		first = False
	uniques.sort(key=lambda u: u[0])
	return [(u[0], u[1], u[2], u[3]) for u in uniques], bad
