"""IL2P codec parameters (reference modems_codecs/il2p.py:109-145); sync search,
Reed-Solomon, descrambling, header translation and the Hamming-protected trailing
CRC run on the GPU (csrc/il2p.cu)."""
from .. import _lib
from .string_ops import check_boolean


class IL2PCodec:
	codec_kind = _lib.PM_CODEC_IL2P

	def __init__(self, **kwargs):                # il2p.py:110-116
		self.collect_trailing_crc = kwargs.get('crc', True)
		self.identifier = kwargs.get('ident', 1)
		self.min_distance = kwargs.get('min_dist', 0)
		self.disable_rs = kwargs.get('disable_rs', False)
		self.sync_tolerance = kwargs.get('sync_tol', 0)

	def StringOptionsRetune(self, options):      # il2p.py:140-145
		self.collect_trailing_crc = check_boolean(options.get('crc', 'yes'))
		self.disable_rs = check_boolean(options.get('disable_rs', 'no'))
		self.min_distance = int(options.get('min_dist', self.min_distance))
		self.sync_tolerance = int(options.get('sync_tol', self.sync_tolerance))

	def decode(self, stream):
		"""il2p.py:360-519 on the GPU: list[AddressedData] -> list of PacketMeta (SourceDecoder = ident), decoded from the
		initial state; CalculatedCRC / CarriedCRC / ValidCRC / ValidHeader are already filled in."""
		from ..engine import addressed_arrays, stage_engine
		data, addresses = addressed_arrays(stream)
		eng = stage_engine(codec=self, name=self.identifier)
		return eng.packets(*eng.decode_stream(0, data, addresses))[0]

	def describe(self, desc):
		desc.codec_kind = self.codec_kind
		desc.il2p_crc = 1 if self.collect_trailing_crc else 0
		desc.il2p_disable_rs = 1 if self.disable_rs else 0
		desc.il2p_min_dist = int(self.min_distance)
		desc.il2p_sync_tol = int(self.sync_tolerance)
