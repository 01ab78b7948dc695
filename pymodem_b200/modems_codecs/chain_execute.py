"""Chain execution on the GPU.  process_chain / multiprocess_chain keep the
signatures of reference modems_codecs/chain_execute.py:6-52; process_chains is
the batched form that replaces the per-chain process fan-out of
pymodem.py:140-166 (one engine call for every chain)."""
from ..engine import Engine, engine_for


def process_chains(demod_stack, input_audio, **engine_options):
	"""[[name, modem, slicer, stream, codec], ...], int16 ndarray ->
	[list[PacketMeta] per chain], chains in the order given."""
	eng = engine_for(demod_stack, **engine_options)
	return eng.run(input_audio)


def process_recordings(demod_stack, recordings, **engine_options):
	"""Many recordings through the same demod_stack in ONE engine call (pm_engine_run_batch):
	[int16 ndarray, ...] -> [per recording [per chain list of PacketMeta]].  The reference would run len(recordings) x
	len(demod_stack) processes one after the other (pymodem.py:140-166 handles one file per invocation); here all of
	them are resident at once, which is what keeps a B200 busy when the modem has a sequential carrier loop."""
	eng = engine_for(demod_stack, recordings=len(recordings), **engine_options)
	return eng.run_batch(recordings)


def process_chain(chain, input_audio):
	return process_chains([chain], input_audio)[0]


def multiprocess_chain(chain, input_audio, queue):
	queue.put(process_chain(chain, input_audio))
	return
