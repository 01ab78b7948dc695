"""Carrier types of the demod_chain path, with the attribute names the reference's blocks exchange
(modems_codecs/data_classes.py:7-17): the slicers return AddressedData, MPSKModem.demod returns IQData."""


class AddressedData:
	"""Sliced bytes and, for each of them, the 1-based index of the soft sample that completed it (slicer.py:75, 92-97).
	Extra positional arguments are accepted and ignored, as in the reference."""

	def __init__(self, data, address, *_ignored):
		self.data, self.address = data, address


class IQData:
	"""In-phase and quadrature soft streams of a quadrature demodulator (psk.py:748-751)."""

	def __init__(self, i_data=None, q_data=None):
		self.i_data = [] if i_data is None else i_data
		self.q_data = [] if q_data is None else q_data
