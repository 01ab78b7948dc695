"""Sharded execution on the GPU: several shards of one recording (emulated as several engines
on one device, run in lock step by the same ShardWorker protocol the multi-GPU path uses)
must reproduce the unsharded / reference result bit for bit."""
import numpy as np
import pytest

from util import Golden, as_tuples

pytestmark = pytest.mark.gpu


def build_stack(sample_rate, lines):
	from pymodem_b200.modems_codecs import chain_builder
	return [chain_builder.build_chain(sample_rate, l) for l in lines if l.get("object_type") == "demod_chain"]


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("tag", ["afsk1200_superopt_48k", "fsk9600_ax25_48k"])
def test_sharded_equals_reference(cuda_lib, tag, world):
	from pymodem_b200.sharded import run_sharded_local
	g = Golden(tag)
	got, info = run_sharded_local(build_stack(g.sample_rate, g.lines), g.audio(), world, tail_bits=2048,
		segment_len=4096, warmup_len=8192)
	assert as_tuples(got) == g.all_packets()


def test_sharded_without_warmup_is_repaired_by_the_handoff(cuda_lib):
	from pymodem_b200.sharded import run_sharded_local
	g = Golden("afsk1200_superopt_48k")
	got, info = run_sharded_local(build_stack(g.sample_rate, g.lines), g.audio(), 4, tail_bits=2048,
		segment_len=4096, warmup_len=0)
	assert as_tuples(got) == g.all_packets()
	assert sum(info['repairs']) > 0


def _with_silence(oracle):
	from pymodem_b200 import configs, synth
	lines = configs.afsk_1200_ax25_super_opt()
	audio = synth.afsk1200_ax25(duration_s=14.0, sample_rate=48000, frame_interval_s=0.8, noise_start=0.05,
		noise_end=0.6, seed=41, noise_seed=42, first_frame_s=0.1)[0].copy()
	audio[48000 * 3:48000 * 9] = 0          # six seconds of digital silence: no zero crossings at all
	return lines, audio, oracle.run_config(48000, lines, audio)


@pytest.mark.parametrize("passes", [0, 1, 6])
def test_silence_stretch_unsharded(cuda_lib, oracle, passes):
	"""No zero crossings -> speculated start states never converge; the verify passes cascade and the
	sequential sweep (verify_passes exhausted) finishes the job exactly."""
	from pymodem_b200.engine import Engine
	lines, audio, want = _with_silence(oracle)
	eng = Engine(build_stack(48000, lines), segment_len=4096, warmup_len=4096, verify_passes=passes)
	try:
		got = as_tuples(eng.run(audio))
		st = eng.stats()
	finally:
		eng.close()
	assert got == want
	assert st["slicer_repairs"] > 0
	assert sum(len(w) for w in want) > 0


def test_silence_stretch_sharded(cuda_lib, oracle):
	from pymodem_b200.sharded import run_sharded_local
	lines, audio, want = _with_silence(oracle)
	got, info = run_sharded_local(build_stack(48000, lines), audio, 4, tail_bits=2048, segment_len=4096, warmup_len=4096)
	assert as_tuples(got) == want
	assert max(info['rounds']) >= 2


def test_frame_straddling_a_boundary(cuda_lib, oracle):
	"""8 s, 2 shards: the boundary (sample 192512 = 4.01 s) falls inside the frame sent at 3.7 s, whose
	data starts ~117 bits before it.  With a 2048-bit tail the frame is decoded by rank 1 exactly; a
	64-bit tail cannot be exact and the engine must say so instead of guessing."""
	from pymodem_b200 import configs, synth
	from pymodem_b200.engine import EngineError
	from pymodem_b200.sharded import run_sharded_local
	lines = configs.afsk_1200_ax25_super_opt()
	audio = synth.afsk1200_ax25(duration_s=8.0, sample_rate=48000, frame_interval_s=0.7, noise_start=0.02,
		noise_end=0.05, seed=51, noise_seed=52, first_frame_s=0.2)[0]
	want = oracle.run_config(48000, lines, audio)
	stack = build_stack(48000, lines)
	got, info = run_sharded_local(stack, audio, 2, tail_bits=2048, segment_len=4096, warmup_len=8192)
	assert info['plans'][1]['sample_base'] + info['plans'][1]['own_begin'] == 192512
	assert as_tuples(got) == want
	assert any(3.7 * 48000 < a < 4.6 * 48000 for a, _d, _c in want[0])      # the straddling frame is in the result
	with pytest.raises(EngineError):
		run_sharded_local(stack, audio, 2, tail_bits=64, segment_len=4096, warmup_len=8192)
