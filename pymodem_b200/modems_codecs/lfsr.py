"""Feed-forward descrambler (reference modems_codecs/lfsr.py:10-52): parameters on the host, the GF(2) FIR on the GPU
(csrc/bits.cu lfsr_kernel).  stream_unscramble_8bit() is the reference's duck-typed stage call
(chain_execute.py:40-43); it starts from an empty shift register on every call (see slicer.py of this package)."""
from .string_ops import check_boolean


class LFSR:
	def __init__(self, **kwargs):
		self.polynomial = kwargs.get('poly', 0x1)
		self.invert = kwargs.get('invert', False)
		self.shift_register = 0

	def StringOptionsRetune(self, options):      # lfsr.py:18-20: poly is a hex string
		self.polynomial = int(options.get('poly', 0x1), 16)
		self.invert = check_boolean(options.get('invert', "false"))

	def stream_unscramble_8bit(self, stream):
		"""lfsr.py:22-52 on the GPU: list[AddressedData] -> list[AddressedData] (addresses pass through)."""
		from ..engine import addressed_arrays, addressed_list, stage_engine
		data, addresses = addressed_arrays(stream)
		return addressed_list(*stage_engine(stream=self).unscramble_stream(0, data, addresses))

	def describe(self, desc):
		if self.polynomial <= 0 or self.polynomial >= (1 << 64):
			raise ValueError("LFSR polynomial must fit 64 bits")
		desc.lfsr_poly = self.polynomial
		desc.lfsr_invert = 1 if self.invert else 0
