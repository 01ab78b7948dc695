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


# ---- shard link: the same hand-off carried out by the GPU(s) over peer memory (csrc/link.cu) --------------------
@pytest.mark.parametrize("world", [1, 2, 3, 4])
@pytest.mark.parametrize("tag", ["afsk1200_superopt_48k", "fsk9600_ax25_48k"])
def test_linked_equals_reference(cuda_lib, tag, world):
	"""Every rank ends up with the merged records of all ranks, identical to the unsharded reference result."""
	from pymodem_b200.sharded import run_linked_local
	g = Golden(tag)
	got, info = run_linked_local(build_stack(g.sample_rate, g.lines), g.audio(), world, tail_bits=2048,
		segment_len=4096, warmup_len=16384)
	assert as_tuples(got) == g.all_packets()
	if all(info['verified']):
		assert info['all_ranks_equal']


def test_linked_falls_back_when_a_handoff_does_not_verify(cuda_lib):
	"""No warm-up: the speculated start states are wrong, all ranks agree on that and the repair protocol finishes."""
	from pymodem_b200.sharded import run_linked_local
	g = Golden("afsk1200_superopt_48k")
	got, info = run_linked_local(build_stack(g.sample_rate, g.lines), g.audio(), 4, tail_bits=2048,
		segment_len=4096, warmup_len=0)
	assert not any(info['verified'])
	assert as_tuples(got) == g.all_packets()


def test_linked_frame_straddling_a_boundary(cuda_lib, oracle):
	from pymodem_b200 import configs, synth
	from pymodem_b200.engine import EngineError
	from pymodem_b200.sharded import run_linked_local
	lines = configs.afsk_1200_ax25_super_opt()
	audio = synth.afsk1200_ax25(duration_s=8.0, sample_rate=48000, frame_interval_s=0.7, noise_start=0.02,
		noise_end=0.05, seed=51, noise_seed=52, first_frame_s=0.2)[0]
	want = oracle.run_config(48000, lines, audio)
	stack = build_stack(48000, lines)
	got, info = run_linked_local(stack, audio, 2, tail_bits=2048, segment_len=4096, warmup_len=16384)
	assert as_tuples(got) == want
	# a hand-off tail far shorter than a frame: the shard that holds the closing flag cannot decode it from what it has --
	# round 1 gave up with an error here; now all shards recover from the gathered bitstream
	got, info = run_linked_local(stack, audio, 2, tail_bits=128, segment_len=4096, warmup_len=16384)
	assert as_tuples(got) == want
	assert info.get('recovered')


def test_linked_engine_is_reusable(cuda_lib):
	"""Epoch flags and the two parity slots: several runs through the same link, alternating inputs."""
	from pymodem_b200.engine import Engine
	from pymodem_b200.sharded import ShardWorker, local_exchange, plan_shards, run_protocol
	g = Golden("afsk1200_superopt_48k")
	stack = build_stack(g.sample_rate, g.lines)
	audio = g.audio()
	half = np.ascontiguousarray(audio[: len(audio) // 2 + 7777])
	world = 2
	engines = [Engine(stack, segment_len=4096, warmup_len=32768) for _ in range(world)]
	try:
		bases = [e.link_create(r, world, 2048, len(audio))[1] for r, e in enumerate(engines)]
		for e in engines:
			e.link_connect(pointers=bases)
		outs, verdicts = [], []
		for rec in (audio, half, audio, audio):
			plans = plan_shards(len(rec), world, segment_len=4096, warm_len=32768, trim_max=305, samples_per_symbol=40.0, tail_bits=2048)
			loc = [np.ascontiguousarray(rec[p['audio_begin']:p['audio_end']]) for p in plans]
			if not outs:                     # size every device buffer up front (see run_linked_local)
				run_protocol([ShardWorker(e, p, l.ctypes.data, len(l)) for e, p, l in zip(engines, plans, loc)], local_exchange)
			for e, p, l in zip(engines, plans, loc):
				e.run_linked_begin(l.ctypes.data, len(l), p)
			verdict = [e.run_linked_end() for e in engines]
			assert all(verdict) or not any(verdict)          # every rank sees the same states
			if all(verdict):
				outs.append([as_tuples(e.packets(*e.fetch())) for e in engines])
			else:                                             # a start state did not verify: repair protocol
				recs, arena = run_protocol([ShardWorker(e, p, l.ctypes.data, len(l)) for e, p, l in zip(engines, plans, loc)],
					local_exchange, resume=True)
				outs.append([as_tuples(engines[0].packets(recs, arena))] * world)
			verdicts.append(all(verdict))
	finally:
		for e in engines:
			e.close()
	assert outs[0][0] == outs[0][1] == g.all_packets()
	assert outs[2] == outs[0] and outs[3] == outs[0]
	assert verdicts[0] == verdicts[2] == verdicts[3]
	assert outs[1][0] == outs[1][1] and outs[1][0] != outs[0][0]


# ---- IL2P chains across shard boundaries: decoder state handed from rank to rank --------------------------------
@pytest.mark.parametrize("tag,world,tail", [("afsk1200_il2p_48k", 2, 4096), ("afsk1200_il2p_48k", 3, 4096),
	("fsk9600_il2p_48k", 2, 4096), ("fsk9600_il2p_48k", 4, 4096)])
def test_sharded_il2p_equals_reference(cuda_lib, tag, world, tail):
	from pymodem_b200.sharded import run_sharded_local
	g = Golden(tag)
	got, info = run_sharded_local(build_stack(g.sample_rate, g.lines), g.audio(), world, tail_bits=tail,
		segment_len=4096, warmup_len=8192)
	assert as_tuples(got) == g.all_packets()


def test_sharded_il2p_noisy_many_boundaries(cuda_lib, oracle):
	"""Heavy noise, frames of every size back to back, 2..6 shards: failed headers and blocks leak their corrected-byte
	counts across boundaries (il2p.py:200-211), frames straddle them, false syncs end right before them."""
	import json
	from pymodem_b200 import synth
	from pymodem_b200.sharded import run_sharded_local
	audio = synth.fsk9600_il2p(duration_s=12.0, sample_rate=48000, frame_interval_s=0.1, noise_start=0.4, noise_end=1.0,
		seed=91, noise_seed=92, first_frame_s=0.02, payload_len=[None, 500, 3, 0, 239, 240, 1023])[0]
	lines = [json.loads(json.dumps(l)) for l in Golden("fsk9600_il2p_48k").chain_lines()]
	for l in lines:
		if l["codec"]["type"] == "il2p":
			l["codec"]["options"]["sync_tol"] = "3"
	want = oracle.run_config(48000, lines, audio)
	assert sum(len(w) for w in want) > 20
	stack = build_stack(48000, lines)
	for world in (2, 3, 5, 6):
		got, info = run_sharded_local(stack, audio, world, tail_bits=12288, segment_len=4096, warmup_len=8192)
		assert as_tuples(got) == want, f"world {world}"


# ---- IL2P chains over the shard link: the decoder state travels from GPU to GPU (csrc/link.cu link_il2p_*) ------
@pytest.mark.parametrize("tag,world,tail", [("afsk1200_il2p_48k", 2, 4096), ("afsk1200_il2p_48k", 3, 4096),
	("fsk9600_il2p_48k", 2, 4096), ("fsk9600_il2p_48k", 4, 4096), ("afsk1200_il2p_48k", 1, 4096)])
def test_linked_il2p_equals_reference(cuda_lib, tag, world, tail):
	"""Mixed AX.25 / IL2P configs through run_linked_begin: where the previous rank's IL2P walk stands arrives through
	the link buffer, not through the host.  Every rank ends with the fixture's packet set (either straight from the
	link, or -- a frame longer than the tail -- from the gathered bitstream)."""
	from pymodem_b200.sharded import run_linked_local
	g = Golden(tag)
	got, info = run_linked_local(build_stack(g.sample_rate, g.lines), g.audio(), world, tail_bits=tail,
		segment_len=4096, warmup_len=16384)
	assert as_tuples(got) == g.all_packets()
	if all(v == 1 for v in info['verified']):
		assert info['all_ranks_equal']


def test_linked_il2p_noisy_many_boundaries(cuda_lib, oracle):
	"""The noisy many-frame recording of test_sharded_il2p_noisy_many_boundaries through the link: leaked corrected-byte
	counts, frames straddling boundaries and false syncs ending right before them all cross inside the link buffer."""
	import json
	from pymodem_b200 import synth
	from pymodem_b200.sharded import run_linked_local
	audio = synth.fsk9600_il2p(duration_s=12.0, sample_rate=48000, frame_interval_s=0.1, noise_start=0.4, noise_end=1.0,
		seed=91, noise_seed=92, first_frame_s=0.02, payload_len=[None, 500, 3, 0, 239, 240, 1023])[0]
	lines = [json.loads(json.dumps(l)) for l in Golden("fsk9600_il2p_48k").chain_lines()]
	for l in lines:
		if l["codec"]["type"] == "il2p":
			l["codec"]["options"]["sync_tol"] = "3"
	want = oracle.run_config(48000, lines, audio)
	stack = build_stack(48000, lines)
	straight = 0
	for world in (2, 3, 5):
		got, info = run_linked_local(stack, audio, world, tail_bits=12288, segment_len=4096, warmup_len=16384)
		assert as_tuples(got) == want, f"world {world}"
		straight += all(v == 1 for v in info['verified'])
	assert straight >= 1          # at least one split decodes through the link itself, without recovery or fall-back
