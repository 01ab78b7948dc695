"""AX.25 frames and flag-free gaps long enough to overflow max_packet_length (ax25.py:46-51: after 1024 bytes byte_index
and -- in the '1' branch -- one_count are reset, which can desynchronise the decoder from the stateless flag pattern).
The GPU path replays such stretches sequentially (csrc/bits.cu ax25_sequential_kernel); sharded runs recover by
decoding the gathered bitstream (sharded.recover_from_bitstream).  Everything is compared with the oracle."""
import numpy as np
import pytest

from util import as_tuples

pytestmark = pytest.mark.gpu


def _stream_from_bits(bits):
	bits = np.asarray(bits, dtype=np.uint8)
	n = len(bits) // 8
	data = np.packbits(bits[:8 * n], bitorder='big')
	addr = (np.arange(n, dtype=np.int64) + 1) * 320
	return data, addr


def _decode_both(oracle, data, addr):
	from pymodem_b200.modems_codecs import chain_builder
	want = oracle.AX25Codec("t").decode(data, addr)
	got = as_tuples([chain_builder.CodecConfigurator({"type": "ax25"}, "t").decode((data, addr))])[0]
	return got, want


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_bits_with_long_gaps(cuda_lib, oracle, seed):
	"""P(1) = 0.3: a flag every ~2800 bits on average, so about one gap in twenty exceeds 8192 bits."""
	rng = np.random.default_rng(seed)
	bits = (rng.random(400000) < 0.3).astype(np.uint8)
	got, want = _decode_both(oracle, *_stream_from_bits(bits))
	assert got == want
	assert any(len(p[1]) > 1023 for p in want) or len(want) > 0


@pytest.mark.parametrize("fill", ["random", "ones", "zeros", "ramp"])
def test_long_frames_back_to_back(cuda_lib, oracle, fill):
	"""Frames of 1000 .. 3000 payload bytes (the 1024th byte completes at every possible phase of the bit stuffing)."""
	from pymodem_b200 import synth
	rng = np.random.default_rng(5)
	bits = []
	for k, plen in enumerate([1000, 1005, 1006, 1007, 1008, 1100, 1101, 1500, 2047, 2048, 3000, 40, 1100]):
		if fill == "random":
			payload = bytes(int(x) for x in rng.integers(0, 256, plen))
		elif fill == "ones":
			payload = bytes([0xFF, 0xFE, 0x7F, 0xFB][k % 4] for _ in range(plen))
		elif fill == "zeros":
			payload = bytes(plen)
		else:
			payload = bytes((i * 7 + k) & 0xFF for i in range(plen))
		frame = synth.ax25_ui_frame("MODEM", "NOISE", payload)
		bits.append(synth.hdlc_bits(frame, preamble_flags=1 + k % 3, postamble_flags=1))
	bits = np.concatenate(bits)
	for shift in range(8):
		got, want = _decode_both(oracle, *_stream_from_bits(np.concatenate([np.zeros(shift, dtype=np.uint8), bits])))
		assert got == want
		assert sum(len(p[1]) > 1023 for p in want) >= 4


def _long_frame_audio():
	from pymodem_b200 import configs, synth
	lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())[:3]
	audio = synth.afsk1200_ax25(duration_s=46.0, sample_rate=48000, frame_interval_s=9.0, noise_start=0.0, noise_end=0.25,
		seed=61, noise_seed=62, first_frame_s=0.3, payload_len=[1100, 60, 1300, 1024, 1500])[0]
	return lines, audio


def test_long_frames_unsharded(cuda_lib, oracle):
	from pymodem_b200.modems_codecs import chain_builder, chain_execute
	lines, audio = _long_frame_audio()
	want = oracle.run_config(48000, lines, audio)
	assert sum(len(p[1]) > 1023 for c in want for p in c) >= 3
	got = as_tuples(chain_execute.process_chains([chain_builder.build_chain(48000, l) for l in lines], audio))
	assert got == want


@pytest.mark.parametrize("world,tail_bits", [(2, 16384), (3, 16384), (3, 2048)])
def test_long_frames_sharded_recover(cuda_lib, oracle, world, tail_bits):
	"""A shard that sees a gap of 8192+ bits (or a frame longer than its hand-off tail) cannot finish from what it
	holds: all shards export their bits and decode the gathered stream -- the result equals the unsharded one."""
	from pymodem_b200.modems_codecs import chain_builder
	from pymodem_b200.sharded import run_linked_local, run_sharded_local
	lines, audio = _long_frame_audio()
	want = oracle.run_config(48000, lines, audio)
	stack = [chain_builder.build_chain(48000, l) for l in lines]
	got, info = run_sharded_local(stack, audio, world, tail_bits=tail_bits)
	assert as_tuples(got) == want
	assert info.get('recovered')
	got, info = run_linked_local(stack, audio, world, tail_bits=tail_bits)
	assert as_tuples(got) == want
	assert info.get('recovered')


def test_il2p_back_to_back_header_only_frames(cuda_lib, oracle):
	"""packet buffers were sized for 152 stream bits per packet; back-to-back IL2P header-only frames
	without trailing CRC take 144.  The buffers grow and the run repeats instead of failing."""
	from pymodem_b200 import synth
	from pymodem_b200.modems_codecs import chain_builder
	frames = []
	for k in range(400):
		air = synth.il2p_frame("MODEM", "NOISE", b"", trailing_crc=False)[0]
		frames.append(synth.il2p_bits(air, preamble_bytes=0, postamble_bytes=0))
	bits = np.concatenate(frames)
	data, addr = _stream_from_bits(bits)
	opts = {"crc": "no", "disable_rs": "no", "min_dist": "0", "sync_tol": "0"}
	want = oracle.IL2PCodec("t", opts).decode(data, addr)
	got = as_tuples([chain_builder.CodecConfigurator({"type": "il2p", "options": opts}, "t").decode((data, addr))])[0]
	assert got == want and len(want) >= 390
