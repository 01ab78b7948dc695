"""AX.25 HDLC codec parameters (reference modems_codecs/ax25.py:11-23); decoding
runs on the GPU (csrc/bits.cu ax25_* kernels)."""
from .. import _lib


class AX25Codec:
	codec_kind = _lib.PM_CODEC_AX25

	def __init__(self, **kwargs):
		self.min_packet_length = kwargs.get('min_packet_length', 18)
		self.max_packet_length = kwargs.get('max_packet_length', 1023)
		self.identifier = kwargs.get('ident', 1)
		if (self.min_packet_length, self.max_packet_length) != (18, 1023):
			raise NotImplementedError("the GPU AX.25 decoder is built for the reference limits 18/1023")

	def decode(self, stream):
		"""ax25.py:25-93 on the GPU: list[AddressedData] -> list of PacketMeta (SourceDecoder = ident), decoded from the
		initial state; CalculatedCRC / CarriedCRC / ValidCRC / ValidHeader are already filled in."""
		from ..engine import addressed_arrays, stage_engine
		data, addresses = addressed_arrays(stream)
		eng = stage_engine(codec=self, name=self.identifier)
		return eng.packets(*eng.decode_stream(0, data, addresses))[0]

	def describe(self, desc):
		desc.codec_kind = self.codec_kind
