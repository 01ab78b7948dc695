"""The FP32 front end's sign guard (csrc/front.cu epilogue), proven instead of swept: over ten minutes each
of four adversarial recordings the packed sign streams of the FP32 route (after the float64 fix-up of the flagged
samples) are compared WORD FOR WORD with the all-float64 route (option precise=1: the reference's own formula,
afsk.py:148-167, in float64) for all 8 chains of the super-opt config.  Only the sign of a soft sample reaches the
slicer (slicer.py:85, 99-102), so equal sign streams are equal results.

The guard flags |y| < 2^-18 (|L_mark| + g |L_space|) + c_abs 2^-24 max|audio of the tile| sum|h_bpf| N_corr sum|h_lpf| (1 + g):
the second term is what the band-pass's rounding at RAW-input magnitude (DC, hum, out-of-band tones) can leave in y.
The test also reports the flag rate and shows that the second term is necessary: with guard_abs = 0 the same
recordings do flip signs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

FS = 48000
SECONDS = 600.0


def _stack():
	from pymodem_b200 import configs
	from pymodem_b200.modems_codecs import chain_builder
	return [chain_builder.build_chain(FS, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]


def _signs(stack, audio, **opts):
	from pymodem_b200.engine import Engine
	eng = Engine(stack, **opts)
	try:
		eng.run_raw(audio)
		st = eng.stats()
		return [eng.signs(c) for c in range(len(stack))], st
	finally:
		eng.close()


def _noise(seed, n, dbfs):
	return np.random.default_rng(seed).normal(0.0, 32767 * 10 ** (dbfs / 20), n)


def _case(name):
	from pymodem_b200 import synth
	n = int(SECONDS * FS)
	t = np.arange(n, dtype=np.float64) / FS
	if name == "noise_under_dc":
		x = _noise(1, n, -50) + 32000.0
	elif name == "noise_under_hum":
		x = _noise(2, n, -50) + 0.6 * 32767 * np.sin(2 * np.pi * 50 * t) + 0.3 * 32767 * np.sin(2 * np.pi * 100 * t)
	elif name == "clipped":
		x = 4.0 * synth.afsk1200_ax25(duration_s=SECONDS, sample_rate=FS, frame_interval_s=2.0, noise_start=0.0, noise_end=1.6,
			seed=71, noise_seed=72)[0].astype(np.float64)
	elif name == "bench_tail":
		# the statistics of the last five minutes of bench.py's hour (AWGN sigma 1.47 .. 1.6 x signal), twice over
		x = synth.afsk1200_ax25(duration_s=SECONDS, sample_rate=FS, frame_interval_s=3.1, noise_start=1.6 * 55 / 60,
			noise_end=1.6, seed=1000, noise_seed=1001)[0].astype(np.float64)
	elif name == "weak_under_dc":
		x = synth.afsk1200_ax25(duration_s=SECONDS, sample_rate=FS, frame_interval_s=2.0, amplitude=0.003, noise_start=0.0,
			noise_end=0.6, seed=73, noise_seed=74)[0].astype(np.float64) + 30000.0
	else:
		raise KeyError(name)
	return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


@pytest.mark.parametrize("name", ["noise_under_dc", "noise_under_hum", "clipped", "bench_tail", "weak_under_dc"])
def test_fp32_signs_equal_float64_route(cuda_lib, name):
	stack = _stack()
	audio = _case(name)
	ref, _ = _signs(stack, audio, precise=1)
	got, st = _signs(stack, audio)
	words = sum(len(r) for r in ref)
	differing = sum(int(np.count_nonzero(a != b)) for a, b in zip(got, ref))
	rate = st["guard_flagged"] / (len(stack) * len(audio))
	print(f"\n[guard] {name}: {words} sign words compared, {differing} differ, guard flag rate {rate:.3e} "
		f"({st['guard_flagged']} samples), fix-up {st['fixup_ms']:.3f} ms")
	assert differing == 0
	assert rate < 0.02          # the guard must not get there by re-evaluating everything


def test_raw_input_term_is_needed(cuda_lib):
	"""Round 1's guard (relative term only) does flip signs under full-scale out-of-band energy."""
	stack = _stack()
	audio = _case("noise_under_hum")[:int(120 * FS)]
	ref, _ = _signs(stack, audio, precise=1)
	old, st = _signs(stack, audio, guard_abs=0.0)
	differing = sum(int(np.count_nonzero(a != b)) for a, b in zip(old, ref))
	print(f"\n[guard] relative term alone: {differing} sign words differ in 120 s x 8 chains")
	assert differing > 0
