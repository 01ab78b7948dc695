"""Carrier types of the demod_chain path (reference modems_codecs/data_classes.py:7-17)."""


class AddressedData:
	def __init__(self, data, address, *args):
		self.data = data
		self.address = address


class IQData:
	def __init__(self):
		self.i_data = []
		self.q_data = []
