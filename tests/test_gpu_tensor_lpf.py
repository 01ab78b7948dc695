"""The AFSK front end's low-pass on the tensor cores (csrc/lpf_tc.cu: three bf16 pieces per operand, six piece
products, FP32 accumulation in TMEM) against the FFMA2 low-pass it replaces and against the reference fixtures: same
packets, same slicer streams, soft values within 1e-5 of the RMS of the reference's float64 values, sign streams equal
to the float64 route's."""
import numpy as np
import pytest

from util import Golden, as_tuples

pytestmark = pytest.mark.gpu


def _stack(g):
	from pymodem_b200.modems_codecs import chain_builder
	return [chain_builder.build_chain(g.sample_rate, l) for l in g.chain_lines()]


@pytest.mark.parametrize("tensor", [1, 0])
@pytest.mark.parametrize("tag", ["afsk1200_superopt_48k", "afsk1200_ax25_44k1", "afsk1200_il2p_48k"])
def test_packets_and_soft_values(cuda_lib, tag, tensor):
	from pymodem_b200.engine import Engine
	g = Golden(tag)
	eng = Engine(_stack(g), tensor_lpf=tensor, keep_soft=1)
	try:
		got = as_tuples(eng.run(g.audio()))
		assert got == g.all_packets()
		for ci in range(g.n_chains):
			if f"c{ci}_soft_dec" not in g.z:
				continue
			soft = eng.soft(ci).astype(np.float64)
			assert len(soft) == int(g.z[f"c{ci}_soft_len"])
			rms = float(g.z[f"c{ci}_soft_rms"])
			err = max(np.max(np.abs(soft[::97] - g.z[f"c{ci}_soft_dec"])), np.max(np.abs(soft[10000:10000 + 8192] - g.z[f"c{ci}_soft_win"])))
			assert err <= 1e-5 * rms, (tag, ci, err / rms)        # tolerance of BASELINE.json's north_star
	finally:
		eng.close()


@pytest.mark.parametrize("n", [1, 305, 306, 337, 4097, 8192 + 305, 8192 + 306, 3 * 8192 + 304, 100001])
def test_ragged_lengths(cuda_lib, oracle, n):
	"""Recording lengths around the low-pass tile (8192 outputs), the front tile and the FIR trim."""
	from pymodem_b200 import configs, synth
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
	audio = synth.afsk1200_ax25(duration_s=3.0, sample_rate=48000, frame_interval_s=0.7, noise_start=0.0, noise_end=0.8,
		seed=91, noise_seed=92, first_frame_s=0.05)[0][:n]
	want = oracle.run_config(48000, lines, audio)
	eng = Engine([chain_builder.build_chain(48000, l) for l in lines], tensor_lpf=1)
	try:
		assert as_tuples(eng.run(audio)) == want
	finally:
		eng.close()


def test_signs_equal_float64_route(cuda_lib):
	from pymodem_b200 import configs, synth
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
	stack = [chain_builder.build_chain(48000, l) for l in lines]
	audio = synth.afsk1200_ax25(duration_s=180.0, sample_rate=48000, frame_interval_s=1.3, noise_start=0.0, noise_end=1.6,
		seed=93, noise_seed=94)[0]
	signs = {}
	for key, opts in (("ref", dict(precise=1)), ("tc", dict(tensor_lpf=1)), ("ffma", dict(tensor_lpf=0))):
		eng = Engine(stack, **opts)
		try:
			eng.run_raw(audio)
			signs[key] = [eng.signs(c) for c in range(len(stack))]
		finally:
			eng.close()
	for key in ("tc", "ffma"):
		assert sum(int(np.count_nonzero(a != b)) for a, b in zip(signs[key], signs["ref"])) == 0, key
