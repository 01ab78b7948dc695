#!/usr/bin/env python
"""bench.py --config NAME: the other shipped configs of BASELINE.json (configs 1-4), each with its own CPU baseline.

  afsk_1200   configs/afsk_1200.json (2 AX.25 + 2 IL2P+CRC chains) on one hour of 48 kHz synthetic AFSK audio
  fsk_9600    configs/fsk_9600.json (IL2P+CRC, IL2P+CRC inverted, G3RUH AX.25) on 10 min of 48 kHz synthetic FSK audio
  bpsk_300    configs/bpsk_300.json (RRC + Costas loop, IL2P+CRC) -- a BATCH of 60 s recordings at 8 kHz
  qpsk_2400   configs/qpsk_2400.json (3 decision-directed MPSK chains, quadrature slicer, IL2P+CRC) -- a batch of 30 s
              recordings at 8 kHz

The carrier-loop modems are sequential per chain (psk.py:173-189, 727-747: NCO wavetable index, phase-error table and
round() feed quantised decisions back, so a segmented loop never becomes bit-identical -- SURVEY Appendix C); one loop
runs at the latency of its float64 dependency chain whatever the GPU.  Throughput comes from running many of them
side by side: pm_engine_run_batch (chain_execute.process_recordings) keeps `--batch` recordings x their chains
resident in one call.  The line reports the per-sample latency of one loop next to the aggregate.

The config lines are read from the committed fixtures (tests/golden/*.npz carry the reference's config files as
shipped), the audio comes from pymodem_b200.synth.  Same JSON contract as bench.py; `metric` is the same
chain-samples/sec.
"""
import json
import os
import statistics
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

UNIT = "chain-samples/s"

WORKLOADS = {
	# name: (fixture holding the config, synth generator, generator arguments, default batch, seconds of the CPU sample)
	"afsk_1200": ("afsk1200_il2p_48k", "afsk1200_il2p", dict(duration_s=3600.0, sample_rate=48000, frame_interval_s=1.0, noise_start=0.0,
		noise_end=1.2, seed=1, noise_seed=2), 1, 30.0),
	"fsk_9600": ("fsk9600_il2p_48k", "fsk9600_il2p", dict(duration_s=600.0, sample_rate=48000, frame_interval_s=0.25, noise_start=0.0,
		noise_end=0.9, seed=3, noise_seed=4), 1, 30.0),
	"bpsk_300": ("bpsk300_il2p_8k", "bpsk300_il2p", dict(duration_s=60.0, sample_rate=8000, frame_interval_s=3.0, noise_start=0.0,
		noise_end=0.9, seed=2004, noise_seed=2005, carrier=1503.0, first_frame_s=1.5, payload_len=[None, 120, 0, 30]), 64, 60.0),
	"qpsk_2400": ("qpsk2400_il2p_8k", "qpsk2400_il2p", dict(duration_s=30.0, sample_rate=8000, frame_interval_s=1.0, noise_start=0.0,
		noise_end=0.7, seed=2006, noise_seed=2007, carrier=1499.0, first_frame_s=1.5, payload_len=[None, 300, 0, 30]), 64, 30.0),
}


def _cpu_worker(args):
	rate, line, audio = args
	from oracle import oracle as orc
	t0 = time.perf_counter()
	n = len(orc.Chain(rate, line).process(audio))
	return time.perf_counter() - t0, n


def cpu_port(rate, lines, recordings, procs):
	"""The oracle port, one process per (recording, chain) work item over `procs` processes -> chain-samples/s."""
	import multiprocessing as mp
	items = [(rate, line, a) for a in recordings for line in lines]
	t0 = time.perf_counter()
	with mp.get_context("fork").Pool(procs) as pool:
		pool.map(_cpu_worker, items)
	dt = time.perf_counter() - t0
	return sum(len(a) for a in recordings) * len(lines) / dt, dt


def main(args):
	import torch
	from util import Golden
	from pymodem_b200 import synth
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	if args.config not in WORKLOADS:
		raise SystemExit(f"bench.py --config: one of super_opt, {', '.join(WORKLOADS)}")
	if int(os.environ.get("WORLD_SIZE", "1")) > 1:
		raise SystemExit("bench.py --config: the extra configs are single-GPU lines")
	fixture, gen, gen_args, batch, cpu_seconds = WORKLOADS[args.config]
	if getattr(args, "batch", 0):
		batch = args.batch
	g = Golden(fixture)
	lines = g.chain_lines()
	rate = gen_args["sample_rate"]
	if args.seconds != 3600.0:
		gen_args = dict(gen_args, duration_s=args.seconds)
	# distinct recordings: the seeds move with the batch index
	recordings = [getattr(synth, gen)(**dict(gen_args, seed=gen_args["seed"] + 10 * r, noise_seed=gen_args["noise_seed"] + 10 * r))[0]
		for r in range(batch)]
	stack = [chain_builder.build_chain(rate, l) for l in lines]
	n_chains = len(stack)
	units = n_chains * sum(len(a) for a in recordings)
	torch.cuda.set_device(0)
	eng = Engine(stack, recordings=batch)

	def step():
		return eng.run_batch(recordings)

	for _ in range(max(args.warmup, 3)):
		out = step()
	torch.cuda.synchronize()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	e0.record()
	t0 = time.perf_counter()
	stats = []
	for _ in range(args.steps):
		out = step()
		stats.append(eng.stats())
	e1.record()
	torch.cuda.synchronize()
	ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
	value = units / (ms * 1e-3)
	n_packets = sum(len(p) for rec in out for p in rec)
	# one recording alone: the latency of one loop per sample (what no amount of batching changes)
	solo = Engine(stack)
	for _ in range(3):
		solo.run(recordings[0])
	t0 = time.perf_counter()
	for _ in range(3):
		solo.run(recordings[0])
	solo_ms = (time.perf_counter() - t0) / 3 * 1e3
	solo_stats = solo.stats()
	solo.close()
	# parity of the batch against single runs is a GPU test (tests/test_gpu_stages.py); here: against the oracle on recording 0
	from oracle import oracle as orc
	orc.build()
	want = orc.run_config(rate, lines, recordings[0])
	got = [[(p.streamaddress, bytes(p.data), p.BytesCorrected) for p in plist] for plist in out[0]]
	cpu = None
	if not args.no_cpu:
		procs = min(8, os.cpu_count() or 1)
		sample = [a[: int(cpu_seconds * rate)] for a in recordings[: max(1, min(batch, procs))]]
		v, dt = cpu_port(rate, lines, sample, procs)
		cpu = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
			"sample": f"{len(sample)} recording(s) x first {len(sample[0]) / rate:g} s x {n_chains} chains, oracle port, one process per "
				f"(recording, chain) over {procs} processes, {dt:.1f} s"}
	loop_kind = lines[0]["modem"]["type"]
	# the carrier loops against their floor: operations on the per-sample dependency chain x the latency of one dependent
	# float64 operation (control -> NCO phase -> index -> wavetable -> mixer -> [rotation, phase table] -> IIR -> PI -> control)
	from pymodem_b200.engine import measure_fp64_chain
	chain_ops = {"bpsk": 17, "afsk_pll": 15, "mpsk": 21}.get(loop_kind)
	loop_floor = None
	if chain_ops:
		ns_op, cyc_op = measure_fp64_chain(0)
		per_sample = solo_stats["front_ms"] * 1e6 / len(recordings[0])
		loop_floor = {"fp64_dependent_op_ns": ns_op, "fp64_dependent_op_cycles": cyc_op, "ops_on_the_chain_per_sample": chain_ops,
			"floor_ns_per_sample": chain_ops * ns_op, "measured_ns_per_sample": per_sample,
			"frac_of_floor": chain_ops * ns_op / per_sample if per_sample else None,
			"note": "measured = the modem stages of one recording alone (FIRs included) per sample; the chain also crosses two "
				"shared-memory table look-ups and integer conversions, which the floor does not count"}
	line = {"metric": "demod chain-samples/sec", "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
		"ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f64" if loop_kind in ("bpsk", "mpsk", "afsk_pll") else "f32", "data": "synthetic",
		"config": {"workload": f"{args.config}.json ({n_chains} chains) x {batch} recording(s) of {len(recordings[0]) / rate:g} s at {rate} Hz, "
			"one pm_engine_run_batch call per step (host buffers: H2D and record D2H inside the timed region)", "chains": n_chains,
			"sample_rate": rate, "batch": batch, "samples_per_recording": len(recordings[0])},
		"e2e": {"value": value, "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": stats[-1]["h2d_bytes"],
			"d2h_bytes_per_step": stats[-1]["d2h_bytes"], "call": "chain_execute.process_recordings / Engine.run_batch"},
		"gpu_launches": sum(s["kernel_launches"] for s in stats),
		"single_recording": {"ms": solo_ms, "value": n_chains * len(recordings[0]) / (solo_ms * 1e-3), "unit": UNIT,
			"ns_per_sample_per_chain": solo_ms * 1e6 / len(recordings[0]),
			"note": "one recording alone: for the carrier-loop modems this is the latency of one thread's float64 dependency chain "
				"(NCO, mixer, IIR, PI: ~25 dependent operations per sample), not a throughput"},
		"loop_latency": loop_floor, "cpu_baseline": cpu, "packets_per_step": n_packets,
		"parity": {"recording_0_equals_oracle": got == want, "n_packets_recording_0": sum(len(p) for p in got)},
		"stage_ms": {k: statistics.mean(s[k] for s in stats) for k in ("total_ms", "front_ms", "fixup_ms", "slicer_ms", "bits_ms", "d2h_ms")},
		"host_cpus": os.cpu_count()}
	print(json.dumps(line), flush=True)
	eng.close()
