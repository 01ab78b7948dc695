"""BASELINE.json's full size in the test suite: the super-opt 8-chain config on the bench's one hour of 48 kHz audio.
The CPU oracle needs 20 s on 8 processes for it (tools/verify_hour.py compares against it directly); here the packet
set is checked through its committed digest (profiles/r02z_verify_hour.txt: GPU == oracle, d794b4007f27f787) and
through size-independent properties: the tensor-core and the FFMA low-pass give the same records, two engines running
concurrently on one GPU give the same records, batched == single, a different slicer geometry gives the same records."""
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ORACLE_HOUR_DIGEST = "d794b4007f27f787"


@pytest.fixture(scope="module")
def hour():
	from pymodem_b200 import configs, synth
	from pymodem_b200.modems_codecs import chain_builder
	audio = synth.afsk1200_ax25(duration_s=3600.0, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
		seed=1000, noise_seed=1001)[0]
	stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]
	return audio, stack


def _digest(per_chain):
	h = hashlib.sha256()
	for ci, plist in enumerate(per_chain):
		for p in plist:
			h.update(repr((ci, int(p.streamaddress), bytes(p.data), int(p.BytesCorrected))).encode())
	return h.hexdigest()[:16]


def _records(eng, audio):
	recs, arena = eng.run_raw(audio)
	return recs.tobytes(), arena.tobytes()


def test_hour_equals_the_oracle_digest(cuda_lib, hour):
	from pymodem_b200.engine import Engine
	audio, stack = hour
	eng = Engine(stack)
	try:
		per_chain = eng.run(audio)
		st = eng.stats()
	finally:
		eng.close()
	assert sum(len(p) for p in per_chain) == 5413
	assert _digest(per_chain) == ORACLE_HOUR_DIGEST
	assert st["slicer_repairs"] < 0.01 * st["slicer_segments"] and st["guard_flagged"] < 1e-4 * 8 * len(audio)


def test_hour_properties(cuda_lib, hour):
	from pymodem_b200.engine import Engine
	audio, stack = hour
	ref_eng = Engine(stack)
	try:
		ref = _records(ref_eng, audio)
		# the low-pass on the FP32 pipe instead of the tensor cores; another slicer geometry
		for opts in (dict(tensor_lpf=0), dict(segment_len=8192, warmup_len=32768, warmup_exact_len=8192)):
			eng = Engine(stack, **opts)
			try:
				assert _records(eng, audio) == ref, opts
			finally:
				eng.close()
		# two recordings in flight on one GPU (what bench.py's e2e.pipelined does): same records from both
		other = Engine(stack)
		out = {}

		def worker(key, e):
			out[key] = [_records(e, audio) for _ in range(2)]
		try:
			threads = [threading.Thread(target=worker, args=(k, e)) for k, e in (("a", ref_eng), ("b", other))]
			for t in threads:
				t.start()
			for t in threads:
				t.join()
		finally:
			other.close()
		assert all(r == ref for rs in out.values() for r in rs)
	finally:
		ref_eng.close()
