"""Feed-forward descrambler parameters (reference modems_codecs/lfsr.py:10-20);
the GF(2) FIR itself runs on the GPU (csrc/bits.cu lfsr_kernel)."""
from .string_ops import check_boolean


class LFSR:
	def __init__(self, **kwargs):
		self.polynomial = kwargs.get('poly', 0x1)
		self.invert = kwargs.get('invert', False)
		self.shift_register = 0

	def StringOptionsRetune(self, options):      # lfsr.py:18-20: poly is a hex string
		self.polynomial = int(options.get('poly', 0x1), 16)
		self.invert = check_boolean(options.get('invert', "false"))

	def describe(self, desc):
		if self.polynomial <= 0 or self.polynomial >= (1 << 64):
			raise ValueError("LFSR polynomial must fit 64 bits")
		desc.lfsr_poly = self.polynomial
		desc.lfsr_invert = 1 if self.invert else 0
