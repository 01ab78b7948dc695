"""The reference's duck-typed stage calls (chain_execute.py:32-47) one at a time on the GPU -- slicer.slice(),
stream.stream_unscramble_8bit(), codec.decode() of the mirror classes -- against the oracle's restatement of the
same blocks fed with the same intermediate data, and the engine cache against retuned blocks."""
import numpy as np
import pytest

from util import Golden, as_tuples

pytestmark = pytest.mark.gpu


def _pairs(stream):
	return [(s.data, s.address) for s in stream]


def _zip(b, a):
	return list(zip(b.tolist(), a.tolist()))


@pytest.mark.parametrize("tag,ci", [("afsk1200_superopt_48k", 0), ("afsk1200_superopt_48k", 5), ("afsk1200_ax25_44k1", 0),
	("fsk9600_ax25_48k", 0)])
def test_binary_slicer_slice(cuda_lib, oracle, tag, ci):
	from pymodem_b200.modems_codecs import chain_builder
	g = Golden(tag)
	line = g.chain_lines()[ci]
	oc = oracle.Chain(g.sample_rate, line)
	soft = oc.modem.demod(g.audio())
	want = oc.slicer.slice(soft)
	chain = chain_builder.build_chain(g.sample_rate, line)
	got = chain[2].slice(soft)
	assert _pairs(got) == _zip(*want) and len(got) > 100
	# the reference's other per-stage methods on the result
	want_u = oc.stream.stream_unscramble_8bit(*want)
	got_u = chain[3].stream_unscramble_8bit(got)
	assert _pairs(got_u) == _zip(*want_u)
	want_p = oc.codec.decode(*want_u)
	got_p = chain[4].decode(got_u)
	assert as_tuples([got_p])[0] == want_p
	assert all(p.SourceDecoder == line['object_name'] for p in got_p)


@pytest.mark.parametrize("tag", ["qpsk2400_il2p_8k", "qpsk2400_il2p_22k"])
def test_quadrature_slicer_slice_and_il2p_decode(cuda_lib, oracle, tag):
	from pymodem_b200.modems_codecs import chain_builder
	from pymodem_b200.modems_codecs.data_classes import IQData
	g = Golden(tag)
	line = g.chain_lines()[1]
	oc = oracle.Chain(g.sample_rate, line)
	i_s, q_s = oc.modem.demod(g.audio())
	want = oc.slicer.slice((i_s, q_s))
	chain = chain_builder.build_chain(g.sample_rate, line)
	got = chain[2].slice(IQData(i_s, q_s))
	assert _pairs(got) == _zip(*want) and len(got) > 100
	want_u = oc.stream.stream_unscramble_8bit(*want)
	got_u = chain[3].stream_unscramble_8bit(got)
	assert _pairs(got_u) == _zip(*want_u)
	want_p = oc.codec.decode(*want_u)
	got_p = chain[4].decode(got_u)
	assert as_tuples([got_p])[0] == want_p and len(want_p) > 0


@pytest.mark.parametrize("poly,invert", [("0x3", "true"), ("0x63003", "true"), ("0x1", "false"), ("0x211", "false")])
def test_lfsr_unscramble(cuda_lib, oracle, poly, invert):
	from pymodem_b200.modems_codecs import chain_builder
	rng = np.random.default_rng(11)
	for n in (0, 1, 3, 4, 5, 1000, 4097):
		data = rng.integers(0, 256, n).astype(np.uint8)
		addr = np.cumsum(rng.integers(280, 360, n)).astype(np.int64)
		opts = {"poly": poly, "invert": invert}
		want = oracle.LFSR(opts).stream_unscramble_8bit(data, addr)
		blk = chain_builder.StreamConfigurator({"type": "lfsr", "options": opts})
		got = blk.stream_unscramble_8bit((data, addr))
		assert _pairs(got) == _zip(*want)
	# SURVEY Appendix B.2 known answers
	kat = chain_builder.StreamConfigurator({"type": "lfsr", "options": {"poly": "0x3", "invert": "true"}})
	from pymodem_b200.modems_codecs.data_classes import AddressedData
	out = kat.stream_unscramble_8bit([AddressedData(b, i + 1) for i, b in enumerate([0x00, 0xFF, 0xAA, 0x0F, 0x7E])])
	assert [s.data for s in out] == [0xFF, 0x7F, 0x80, 0xF7, 0x3E]


def test_empty_inputs(cuda_lib):
	from pymodem_b200.modems_codecs import chain_builder
	assert chain_builder.SlicerConfigurator(48000, {"type": "binary", "config": "1200", "options": {}}).slice(np.zeros(0)) == []
	assert chain_builder.CodecConfigurator({"type": "ax25"}, "x").decode([]) == []
	assert chain_builder.CodecConfigurator({"type": "il2p", "options": {}}, "x").decode([]) == []


def test_engine_cache_follows_retuned_blocks(cuda_lib, oracle):
	"""engine_for() used to key on object identities; a retune between two process_chains calls must take
	effect, as it does in the reference (the blocks' state is read on every call)."""
	from pymodem_b200.modems_codecs import chain_builder, chain_execute
	g = Golden("afsk1200_superopt_48k")
	audio = g.audio()
	lines = g.chain_lines()[2:4]
	stack = [chain_builder.build_chain(g.sample_rate, l) for l in lines]
	first = as_tuples(chain_execute.process_chains(stack, audio))
	assert first == [g.packets(2), g.packets(3)]
	import copy
	changed = copy.deepcopy(lines)
	changed[0]['modem']['options']['space_gain'] = "2.75"
	changed[0]['slicer']['options']['lock_rate'] = "0.6"
	stack[0][1].StringOptionsRetune(changed[0]['modem']['options'])
	stack[0][2].StringOptionsRetune(changed[0]['slicer']['options'])
	second = as_tuples(chain_execute.process_chains(stack, audio))
	assert second == oracle.run_config(g.sample_rate, changed, audio)
	assert second[0] != first[0] and second[1] == first[1]


def test_pageable_and_pinned_input_agree(cuda_lib):
	"""pm_engine_run takes any host pointer: pageable memory goes through the engine's pinned ring, registered memory
	is read in place; the chunked copy must not change a bit."""
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	g = Golden("afsk1200_superopt_48k")
	audio = g.audio()
	stack = [chain_builder.build_chain(g.sample_rate, l) for l in g.chain_lines()]
	for opts in ({}, {"h2d_chunk": 1 << 16, "copy_threads": 1}, {"h2d_chunk": 100000, "copy_threads": 3}):
		eng = Engine(stack, **opts)
		try:
			a = as_tuples(eng.run(audio))
			assert eng.pin(audio)
			b = as_tuples(eng.run(audio))
		finally:
			eng.close()
		assert a == b == g.all_packets()


@pytest.mark.parametrize("tag", ["afsk1200_superopt_48k", "afsk1200_ax25_44k1", "fsk9600_il2p_48k"])
def test_early_tail_of_host_runs(cuda_lib, tag):
	"""Host-buffer runs launch the guard fix-up chunk by chunk while the copy is still going (option early_tail = 1, the
	default) and, with early_tail = 2, the slicer's segments too: same records as with the whole tail after the last chunk
	(0), as the device-resident path and as the fixture -- with many small chunks, short segments (many batches), a guard
	list that overflows and is grown, and with the tensor-core low-pass off."""
	import torch
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	g = Golden(tag)
	audio = g.audio()
	stack = [chain_builder.build_chain(g.sample_rate, l) for l in g.chain_lines()]
	want = g.all_packets()
	dev = torch.from_numpy(audio).cuda()
	for opts in ({"early_tail": 0}, {"early_tail": 1}, {"early_tail": 2}, {"h2d_chunk": 1 << 16}, {"h2d_chunk": 1 << 16, "early_tail": 2},
			{"h2d_chunk": 1 << 17, "segment_len": 4096, "warmup_len": 16384, "early_tail": 2, "early_batches": 5},
			{"h2d_chunk": 1 << 16, "tensor_lpf": 0}, {"h2d_chunk": 1 << 16, "tensor_lpf": 0, "early_tail": 2},
			{"h2d_chunk": 1 << 16, "guard_cap": 1024, "guard_eps": 2.0 ** -9}, {"h2d_chunk": 1 << 16, "guard_cap": 1024, "guard_eps": 2.0 ** -9, "early_tail": 2}):
		eng = Engine(stack, **opts)
		try:
			assert eng.pin(audio)
			for _ in range(2):
				assert as_tuples(eng.run(audio)) == want, opts
			st = eng.stats()
			eng.run_device_ptr(dev.data_ptr(), len(audio))
			assert as_tuples(eng.packets(*eng.fetch())) == want, opts
		finally:
			eng.close()
		if opts.get("guard_cap"):
			assert st["guard_flagged"] > 1024      # the list did overflow and the run was repeated with a longer one


@pytest.mark.parametrize("tag,count", [("bpsk300_il2p_8k", 5), ("qpsk2400_il2p_8k", 3), ("afsk1200_superopt_48k", 3), ("afsk300_full_8k", 2),
	("afsk1200_il2p_48k", 1), ("bpsk1200_il2p_12k", 1)])
def test_batched_recordings_equal_single_runs(cuda_lib, oracle, tag, count):
	"""pm_engine_run_batch: several recordings of different lengths x all chains in one call == one call per recording
	(and == the fixture for the unmodified one)."""
	from pymodem_b200.modems_codecs import chain_builder, chain_execute
	g = Golden(tag)
	audio = g.audio()
	stack = [chain_builder.build_chain(g.sample_rate, l) for l in g.chain_lines()]
	rng = np.random.default_rng(5)
	recs = [audio]
	for k in range(1, count):
		cut = audio[int(rng.integers(0, len(audio) // 3)): len(audio) - int(rng.integers(0, len(audio) // 3))].copy()
		if k % 2:
			cut = np.clip(cut.astype(np.int32) + rng.normal(0, 400, len(cut)).astype(np.int32), -32768, 32767).astype(np.int16)
		recs.append(cut)
	got = chain_execute.process_recordings(stack, recs)
	assert as_tuples(got[0]) == g.all_packets()
	for r in range(1, count):
		assert as_tuples(got[r]) == oracle.run_config(g.sample_rate, g.chain_lines(), recs[r]), r


@pytest.mark.parametrize("rate,config", [(48000, '1200'), (44100, '1200'), (48000, '9600'), (48000, '300'), (48000, '4800'), (22050, '1200')])
def test_slicer_long_stretches_without_crossings(cuda_lib, oracle, rate, config):
	"""Digital silence / a steady tone: stretches of 10^5 samples without a zero crossing.  No speculated start state
	converges there, every segment in them is repaired from its predecessor -- and where samples per symbol is a dyadic
	rational (40, 5, 160, 10; not 36.75 within eight words, not 18.375) the repair copies the exactly repeating clock
	and mask words instead of stepping (SlicerChain::quiet_words).  Bytes and addresses equal the oracle's loop."""
	from pymodem_b200.modems_codecs import slicer as slicer_mod
	rng = np.random.default_rng(rate + len(config))
	parts = []
	for k in range(7):
		parts.append(rng.normal(0.0, 1.0, int(rng.integers(3000, 60000))))                 # crossings at every scale
		level = [1.0, -1.0, 0.0, -0.25][k % 4]
		parts.append(np.full(int(rng.integers(40000, 260000)), level))                      # none at all (0.0 counts as >= 0)
	soft = np.concatenate(parts + [rng.normal(0.0, 1.0, 5000)])
	want = oracle.BinarySlicer(rate, config, {}).slice(soft)
	got = slicer_mod.BinarySlicer(sample_rate=rate, config=config).slice(soft)
	assert _pairs(got) == _zip(*want) and len(got) > 100
