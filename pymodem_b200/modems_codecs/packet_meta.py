"""Decoded-packet records and the streamaddress-based duplicate/unique
correlation (reference modems_codecs/packet_meta.py:178-271).

PacketMeta keeps the reference's field names.  Records that come back from the
GPU engine already carry CalculatedCRC / CarriedCRC / ValidCRC / ValidHeader
(computed by csrc/bits.cu packet_copy_kernel); CalcCRC()/Validate() recompute
them on the host for packets built any other way."""
from collections import Counter


def _crc16_x25(data):
	"""crc_functions.py:44-54"""
	crc = 0xFFFF
	for byte in data:
		crc ^= int(byte)
		for _ in range(8):
			crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
	return crc ^ 0xFFFF


def ValidateHeader(frame):
	"""packet_meta.py:21-41: more than 15 bytes and bytes 0..6 (>>1) are 0 or
	printable ASCII (the reference's subfield counter never resets, so only the
	first seven bytes are tested)."""
	if len(frame) <= 15:
		return False
	for index in range(7):
		c = int(frame[index]) >> 1
		if (c < 32 or c > 126) and c != 0:
			return False
	return True


class ReportStyle:
	def __init__(self, options):
		self.destination = options.get('destination', 'std_out')
		self.style = options.get('style', 'raw')


class PacketMeta:
	def __init__(self):
		self.data = []
		self.streamaddress = 0
		self.source_sample_rate = 0.0
		self.CalculatedCRC = 0
		self.CarriedCRC = 0
		self.ValidCRC = False
		self.SourceDecoder = 0
		self.BytesCorrected = 0
		self.CorrelatedDecoders = []
		self.SlicedIQSamples = []

	def CalcCRC(self):                           # packet_meta.py:197-203
		self.CarriedCRC = int(self.data[-1]) * 256 + int(self.data[-2])
		self.CalculatedCRC = _crc16_x25(self.data[:-2])
		self.ValidCRC = self.CarriedCRC == self.CalculatedCRC
		return self.ValidCRC

	def Validate(self):                          # packet_meta.py:205-208
		self.ValidHeader = ValidateHeader(self.data)


class PacketMetaArray:
	def __init__(self):
		self.raw_packet_arrays = []
		self.unique_packet_array = []

	def add(self, array):
		self.raw_packet_arrays.append(array)

	def CalcCRCs(self):                          # packet_meta.py:219-223
		for array in self.raw_packet_arrays:
			for packet in array:
				if not getattr(packet, '_device_checked', False):
					packet.CalcCRC()
					packet.Validate()

	def Correlate(self, **kwargs):
		"""packet_meta.py:230-271, same result, with the unique list indexed by
		CalculatedCRC so a packet is only compared with same-CRC candidates
		(still in insertion order, so the same first match wins)."""
		self.address_distance = kwargs.get('address_distance', 1000)
		by_crc = {}
		first_array = True
		for raw_packet_array in self.raw_packet_arrays:
			for raw_packet in raw_packet_array:
				if not (raw_packet.ValidCRC and raw_packet.ValidHeader):
					continue
				is_unique = True
				if not first_array:
					for unique_packet in by_crc.get(raw_packet.CalculatedCRC, ()):
						if (unique_packet.SourceDecoder != raw_packet.SourceDecoder
								and abs(raw_packet.streamaddress - unique_packet.streamaddress) < self.address_distance):
							is_unique = False
							unique_packet.CorrelatedDecoders.append(raw_packet.SourceDecoder)
							break
				if is_unique:
					raw_packet.CorrelatedDecoders.append(raw_packet.SourceDecoder)
					self.unique_packet_array.append(raw_packet)
					by_crc.setdefault(raw_packet.CalculatedCRC, []).append(raw_packet)
			first_array = False
		self.unique_packet_array = sorted(self.unique_packet_array, key=lambda packet: packet.streamaddress)
		decoder_unique_list = []
		decoder_list = []
		for packet in self.unique_packet_array:
			decoder_list.extend(packet.CorrelatedDecoders)
			if len(packet.CorrelatedDecoders) == 1:
				decoder_unique_list.append(packet.SourceDecoder)
		self.DecoderUniqueHistogram = Counter(decoder_unique_list)
		self.DecoderHistogram = Counter(decoder_list)

	def CountBad(self):
		self.bad_count = sum(1 for arr in self.raw_packet_arrays for p in arr
			if (p.ValidCRC is False) or (p.ValidHeader is False))
		return self.bad_count

	def CountGood(self):
		self.good_count = sum(1 for p in self.unique_packet_array if p.ValidCRC and p.ValidHeader)
		return self.good_count

	def Report(self, order):
		"""A compact text report (the reference's long-form header pretty printer,
		packet_meta.py:43-169/337-370, is host string formatting outside the
		accelerated path)."""
		lines = []
		count = 0
		for packet in self.unique_packet_array:
			if packet.ValidCRC and packet.ValidHeader:
				count += 1
				text = ''.join(chr(b) if 0x1F < b < 0x7F else f'<{hex(b)}>' for b in (int(x) for x in packet.data[:-2]))
				lines.append(f"Packet number: {count} CRC: {hex(packet.CalculatedCRC)} stream address: {packet.streamaddress}")
				lines.append(f"Source decoders: {packet.CorrelatedDecoders}")
				lines.append(f"Packet byte count: {len(packet.data)} Bytes corrected: {packet.BytesCorrected}")
				lines.append(text)
		lines.append(f"Unique, valid packets: {self.CountGood()}")
		lines.append(f"Packets rejected from all decoders for CRC failure: {self.CountBad()}")
		if hasattr(self, 'DecoderHistogram'):
			lines.append("Total packets by decoder:")
			lines += [f"{d} {c}" for d, c in self.DecoderHistogram.most_common()]
			lines.append("Unique packets by decoder:")
			lines += [f"{d} {c}" for d, c in self.DecoderUniqueHistogram.most_common()]
		return "\n".join(lines)
