"""Symbol-timing slicers -- parameters on the host, the timing loop on the GPU (csrc/slicer.cu + the gather kernels
of csrc/bits.cu).  Mirrors reference modems_codecs/slicer.py (BinarySlicer 9-107, QuadratureSlicer 109-242):
constructor kwargs, presets, StringOptionsRetune, tune(), slice().

slice() is the reference's duck-typed stage call (chain_execute.py:36-39).  One difference, shared by the other
per-stage methods of this package: a call processes a complete recording from the state tune() sets (the reference's
blocks carry their state from one call to the next, slicer.py:50-56; chain_execute.py calls each block once)."""
from .. import _lib


def _slice(slicer, soft_i, soft_q=None):
	from ..engine import addressed_list, stage_engine
	data, addresses = stage_engine(slicer=slicer).slice_soft(0, soft_i, soft_q)
	return addressed_list(data, addresses)

_QPSK_DEMAP = [3, 1, 2, 0, 2, 3, 0, 1, 1, 0, 3, 2, 0, 2, 1, 3]
_BPSK_DEMAP = [0, 0, 1, 1]


class BinarySlicer:
	slicer_kind = _lib.PM_SLICER_BINARY

	def __init__(self, **kwargs):
		self.definition = kwargs.get('config', '1200')
		self.sample_rate = kwargs.get('sample_rate', '8000')
		# slicer.py:22-33
		self.symbol_rate, self.lock_rate = {
			'300': (300, 0.75), '9600': (9600, 0.88), '4800': (4800, 0.88),
		}.get(self.definition, (1200, 0.75))
		self.tune()

	def retune(self, **kwargs):
		self.symbol_rate = kwargs.get('symbol_rate', self.symbol_rate)
		self.lock_rate = kwargs.get('lock_rate', self.lock_rate)
		self.sample_rate = kwargs.get('sample_rate', self.sample_rate)
		self.tune()

	def StringOptionsRetune(self, options):      # slicer.py:43-47
		self.symbol_rate = options.get('symbol_rate', self.symbol_rate)
		self.sample_rate = options.get('sample_rate', self.sample_rate)
		self.lock_rate = float(options.get('lock_rate', self.lock_rate))
		self.tune()

	def tune(self):                              # slicer.py:49-56
		self.phase_clock = 0.0
		self.samples_per_symbol = self.sample_rate / self.symbol_rate
		self.rollover_threshold = (self.samples_per_symbol / 2.0) - 0.5
		self.working_byte = 0
		self.working_bit_count = 0
		self.last_sample = 0.0
		self.streamaddress = 0

	def slice(self, samples):
		"""slicer.py:59-107 on the GPU: float64 soft values -> list[AddressedData]."""
		return _slice(self, samples)

	def describe(self, desc):
		desc.slicer_kind = self.slicer_kind
		desc.slicer_sample_rate = float(self.sample_rate)
		desc.symbol_rate = float(self.symbol_rate)
		desc.lock_rate = float(self.lock_rate)
		desc.state_mask = 0
		desc.bits_per_symbol = 1


class QuadratureSlicer:
	slicer_kind = _lib.PM_SLICER_QUADRATURE

	def __init__(self, **kwargs):
		self.sample_rate = kwargs.get('sample_rate', '8000')
		self.definition = kwargs.get('config', '600')
		# slicer.py:124-165: (state_mask, bits_per_symbol, demap, symbol_rate, lock_rate)
		presets = {
			'qpsk_600': (0xF, 2, _QPSK_DEMAP, 300, 0.815),
			'bpsk_300': (0x3, 1, _BPSK_DEMAP, 300, 0.815),
			'bpsk_1200': (0x3, 1, _BPSK_DEMAP, 1200, 0.9),
			'qpsk_2400': (0xF, 2, _QPSK_DEMAP, 1200, 0.9),
			'qpsk_4800': (0xF, 2, _QPSK_DEMAP, 2400, 0.99),
			'qpsk_3600': (0xF, 2, _QPSK_DEMAP, 1800, 0.99),
		}
		(self.state_mask, self.bits_per_symbol, demap, self.symbol_rate,
			self.lock_rate) = presets.get(self.definition, (0xF, 2, _QPSK_DEMAP, 1200, 0.9))
		self.demap = list(demap)
		self.tune()

	def retune(self, **kwargs):
		self.symbol_rate = kwargs.get('symbol_rate', self.symbol_rate)
		self.lock_rate = kwargs.get('lock_rate', self.lock_rate)
		self.sample_rate = kwargs.get('sample_rate', self.sample_rate)
		self.tune()

	def StringOptionsRetune(self, options):      # slicer.py:176-180
		self.symbol_rate = options.get('symbol_rate', self.symbol_rate)
		self.sample_rate = options.get('sample_rate', self.sample_rate)
		self.lock_rate = float(options.get('lock_rate', self.lock_rate))
		self.tune()

	def tune(self):                              # slicer.py:181-191
		self.phase_clock = 0.0
		self.samples_per_symbol = self.sample_rate / self.symbol_rate
		self.rollover_threshold = (self.samples_per_symbol / 2.0) - 0.5
		self.working_byte = 0
		self.working_bit_count = 0
		self.last_i_sample = 0.0
		self.last_q_sample = 0.0
		self.streamaddress = 0
		self.state_register = 0

	def slice(self, samples):
		"""slicer.py:193-242 on the GPU: IQData (i_data, q_data) -> list[AddressedData]."""
		return _slice(self, samples.i_data, samples.q_data)

	def describe(self, desc):
		desc.slicer_kind = self.slicer_kind
		desc.slicer_sample_rate = float(self.sample_rate)
		desc.symbol_rate = float(self.symbol_rate)
		desc.lock_rate = float(self.lock_rate)
		desc.state_mask = self.state_mask
		desc.bits_per_symbol = self.bits_per_symbol
		for i, v in enumerate(self.demap):
			desc.demap[i] = v
