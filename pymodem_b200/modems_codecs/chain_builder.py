"""String-typed factories for the four block kinds -- same names, arguments and
behaviour as reference modems_codecs/chain_builder.py:17-69 (an unknown or
missing 'type' yields [], as there)."""
from . import afsk, afsk_pll, ax25, fsk, il2p, lfsr, psk, slicer


def ModemConfigurator(arg_sample_rate, input_args):
	new_object = []
	kind = input_args.get('type')
	cls = {'afsk': afsk.AFSKModem, 'fsk': fsk.FSKModem, 'bpsk': psk.BPSKModem, 'mpsk': psk.MPSKModem,
		'afsk_pll': afsk_pll.AFSKPLLModem}.get(kind)
	if kind == 'qpsk':
		# psk.py:197 QPSKModem: the legacy branch-filter Costas loop, used by no shipped config
		raise NotImplementedError("modem type 'qpsk' (legacy QPSKModem) has no GPU path; use 'mpsk'")
	if cls:
		new_object = cls(sample_rate=arg_sample_rate, config=input_args['config'])
		new_object.StringOptionsRetune(input_args['options'])
	return new_object


def SlicerConfigurator(arg_sample_rate, input_args):
	new_object = []
	kind = input_args.get('type')
	cls = {'quadrature': slicer.QuadratureSlicer, 'binary': slicer.BinarySlicer}.get(kind)
	if kind == '4level':
		# broken in the reference as well (slicer.py:312, 432 NameError)
		raise NotImplementedError("4level slicer has no GPU path")
	if cls:
		new_object = cls(sample_rate=arg_sample_rate, config=input_args['config'])
		new_object.StringOptionsRetune(input_args['options'])
	return new_object


def StreamConfigurator(input_args):
	new_object = []
	if input_args.get('type') == 'lfsr':
		new_object = lfsr.LFSR()
		new_object.StringOptionsRetune(input_args['options'])
	return new_object


def CodecConfigurator(input_args, name):
	new_object = []
	kind = input_args['type'].lower()
	if kind == 'il2p':
		new_object = il2p.IL2PCodec(ident=name)
		new_object.StringOptionsRetune(input_args['options'])
	elif kind == 'ax25':
		new_object = ax25.AX25Codec(ident=name)
	return new_object


def build_chain(sample_rate, line):
	"""One demod_chain config line -> [object_name, modem, slicer, stream, codec]
	exactly as pymodem.py:68-114 assembles it."""
	modem = ModemConfigurator(sample_rate, line['modem'])
	slicer_rate = getattr(modem, 'output_sample_rate', sample_rate)    # pymodem.py:86-90
	return [line['object_name'], modem, SlicerConfigurator(slicer_rate, line['slicer']),
		StreamConfigurator(line['stream']), CodecConfigurator(line['codec'], line['object_name'])]
