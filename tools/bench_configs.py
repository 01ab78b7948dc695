"""Throughput of every shipped config family the fixtures cover (BASELINE.json configs 1-4 next to the headline
config 5): GPU engine (host buffers, whole C-ABI call) vs the oracle port on one host core, on synthetic audio.
Config lines come from the committed fixtures (tests/golden), so this runs on the GPU box without the reference."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
from util import Golden
from pymodem_b200 import synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from oracle import oracle as orc

CASES = [
	("afsk_1200.json (2 AX.25 + 2 IL2P chains)", "afsk1200_il2p_48k", lambda s: synth.afsk1200_il2p(s, sample_rate=48000, frame_interval_s=1.0, noise_end=1.2, seed=1, noise_seed=2)[0], 300.0, 20.0),
	("fsk_9600.json (2 IL2P + 1 G3RUH AX.25)", "fsk9600_il2p_48k", lambda s: synth.fsk9600_il2p(s, sample_rate=48000, frame_interval_s=0.25, noise_end=0.9, seed=3, noise_seed=4)[0], 300.0, 20.0),
	("bpsk_300.json (Costas loop, IL2P)", "bpsk300_il2p_8k", lambda s: synth.bpsk300_il2p(s, carrier=1503.0, noise_end=0.9, seed=5, noise_seed=6)[0], 300.0, 60.0),
	("qpsk_2400.json (3 MPSK chains, IL2P)", "qpsk2400_il2p_8k", lambda s: synth.qpsk2400_il2p(s, carrier=1499.0, noise_end=0.7, seed=7, noise_seed=8)[0], 300.0, 30.0),
	("afsk_300.json (2 PLL + 3 correlator chains)", "afsk300_full_8k", None, 80.0, 80.0),
]
print(f"{'config':48s} {'chains':>6s} {'audio s':>8s} {'GPU ms':>9s} {'GPU Mcs/s':>10s} {'CPU s':>7s} {'CPU Mcs/s':>10s} {'ratio':>7s} {'packets':>8s}")
for name, tag, gen, secs, cpu_secs in CASES:
	g = Golden(tag)
	lines = g.chain_lines()
	audio = g.audio() if gen is None else gen(secs)
	stack = [chain_builder.build_chain(g.sample_rate, l) for l in lines]
	eng = Engine(stack)
	for _ in range(2):
		eng.run_raw(audio)
	t0 = time.perf_counter()
	reps = 3
	for _ in range(reps):
		recs, _a = eng.run_raw(audio)
	gpu_s = (time.perf_counter() - t0) / reps
	eng.close()
	sample = audio[: int(cpu_secs * g.sample_rate)]
	t0 = time.perf_counter()
	want = orc.run_config(g.sample_rate, lines, sample)
	cpu_s = time.perf_counter() - t0
	gpu_rate = len(lines) * len(audio) / gpu_s / 1e6
	cpu_rate = len(lines) * len(sample) / cpu_s / 1e6
	print(f"{name:48s} {len(lines):6d} {len(audio) / g.sample_rate:8.0f} {gpu_s * 1e3:9.2f} {gpu_rate:10.1f} {cpu_s:7.2f} {cpu_rate:10.2f} {gpu_rate / cpu_rate:7.0f} {len(recs):8d}", flush=True)
