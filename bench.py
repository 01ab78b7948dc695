#!/usr/bin/env python
"""bench.py -- demod chain-samples/sec of the many-chain AFSK 1200 super-opt config
(reference configs/afsk_1200_ax25_super_opt.json, 8 chains) on one hour of 48 kHz
synthetic AWGN AX.25 audio per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scaling weak|strong]
                  [--seconds S] [--config super_opt|afsk_1200|fsk_9600|bpsk_300|qpsk_2400]

A step = one pass of the whole hot path (FIR front end -> symbol-timing slicer ->
NRZI/LFSR -> AX.25 HDLC + CRC -> packet records on the host) over one batch =
`seconds` of audio x 8 chains on every rank (--scaling strong: `seconds` in total, split over the ranks).
`value` times it with the audio already resident in HBM (pm_engine_run_device).
`e2e` times what a user of the package calls -- chain_execute.process_chains(demod_stack, audio) with the audio in
HOST memory: pm_engine_run (chunked H2D overlapped with the front-end kernels), records D2H, and the per-chain
packet sequences built from them.  Prints ONE JSON line on rank 0.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
	sys.path.insert(0, REPO)

SAMPLE_RATE = 48000
METRIC = "demod chain-samples/sec"
UNIT = "chain-samples/s"
# reference-equivalent (unshared) FP32 work of the AFSK front end
REF_FLOP_PER_CHAIN_SAMPLE = 956.0
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 74.4: 148 SMs x 128 lanes x 2 flop x 1.965 GHz
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12       # 37.2: 64 FP64 lanes per SM
SLICER_STEP_PEAK_G = 2544.5      # G exact slicer steps per second, measured: tools/ubench/slicer_step.cu variant 0, 24 warps per SM
# profiles/r01g_verify_hour.txt: the oracle's packet set for the default hour (seeds 1000/1001), tools/verify_hour.py digest
ORACLE_HOUR_DIGEST = "d794b4007f27f787"


def make_audio(seconds, rank):
	"""The workload of SURVEY.md 8(d): AX.25 UI frames every 3.1 s, Bell-202 AFSK, AWGN sigma
	ramped 0 -> 1.6 x signal amplitude across the recording; seeds differ per rank."""
	from pymodem_b200 import synth
	return synth.afsk1200_ax25(duration_s=seconds, sample_rate=SAMPLE_RATE, frame_interval_s=3.1,
		noise_start=0.0, noise_end=1.6, seed=1000 + 2 * rank, noise_seed=1001 + 2 * rank)[0]


class ClockSampler:
	"""SM clock, power and clock-event (throttle) reasons DURING the timed region (B200_PROFILING.md's clocks line).
	Read in-process through NVML (the library nvidia-smi itself queries) every 25 ms: a looping `nvidia-smi -lms`
	child was measured to stall the driver for milliseconds per poll and inflated a 14 ms step to 16-27 ms.
	Falls back to one-shot nvidia-smi calls when pynvml is missing."""
	SMI_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
		"clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

	def __init__(self, gpu_index, period_s=0.025):
		self.gpu, self.period = gpu_index, period_s
		self.sm, self.power, self.reasons = [], [], set()
		self.sm_max = None
		self.stop_flag = threading.Event()
		self.thread = None
		self.source = None

	def _nvml_loop(self):
		import pynvml as nv
		h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
		self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
		names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
			nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
		while not self.stop_flag.is_set():
			self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
			self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
			bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
			for bit, name in names.items():
				if bits & bit:
					self.reasons.add(name)
			self.stop_flag.wait(self.period)

	def _smi_loop(self):
		while not self.stop_flag.is_set():
			try:
				out = subprocess.run(["nvidia-smi", f"--query-gpu={self.SMI_Q}", "--format=csv,noheader,nounits", "-i",
					str(self.gpu)], capture_output=True, text=True, timeout=10).stdout
				f = [x.strip() for x in out.strip().split(",")]
				self.sm.append(float(f[1])); self.sm_max = float(f[2]); self.power.append(float(f[3]))
				for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
					if val.lower().startswith("active"):
						self.reasons.add(name)
			except Exception:
				pass
			self.stop_flag.wait(0.5)

	def start(self):
		try:
			import pynvml as nv
			nv.nvmlInit()
			# CUDA_VISIBLE_DEVICES may renumber devices: NVML wants the physical index
			vis = os.environ.get("CUDA_VISIBLE_DEVICES")
			if vis:
				ids = [v for v in vis.split(",") if v.strip() != ""]
				if self.gpu < len(ids) and ids[self.gpu].strip().isdigit():
					self.gpu = int(ids[self.gpu])
			target, self.source = self._nvml_loop, "nvml"
		except Exception:
			target, self.source = self._smi_loop, "nvidia-smi"
		self.thread = threading.Thread(target=target, daemon=True)
		self.thread.start()

	def stop(self):
		self.stop_flag.set()
		if self.thread:
			self.thread.join(timeout=15)
		return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.sm_max,
			"power_w_max": max(self.power) if self.power else None, "samples": len(self.sm),
			"reasons": sorted(self.reasons), "source": self.source}


def measured_peaks():
	try:
		with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
			return json.load(f), "measured (MEASURED_PEAKS.json)"
	except (OSError, ValueError):
		return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def records_digest(recs, arena):
	"""sha256 over the packet records in order -- chain, streamaddress, length, BytesCorrected, CRC fields, validity
	flags and the payload bytes -- independent of where the bytes sit in the arena (a merged multi-rank arena is laid
	out rank by rank)."""
	h = hashlib.sha256()
	for f in ("chain", "len", "streamaddress", "bytes_corrected", "calculated_crc", "carried_crc", "valid_crc", "valid_header"):
		h.update(np.ascontiguousarray(recs[f]).tobytes())
	lens = recs["len"].astype(np.int64)
	if len(lens) and lens.sum():
		ends = np.cumsum(lens)
		idx = np.arange(ends[-1], dtype=np.int64) - np.repeat(ends - lens, lens) + np.repeat(recs["offset"].astype(np.int64), lens)
		h.update(np.ascontiguousarray(arena[idx]).tobytes())
	return h.hexdigest()[:16]


def packets_digest(per_chain):
	"""tools/verify_hour.py's digest of a per-chain packet set (the form the committed oracle digest has)."""
	h = hashlib.sha256()
	for ci, plist in enumerate(per_chain):
		for p in plist:
			h.update(repr((ci, int(p.streamaddress), bytes(p.data), int(p.BytesCorrected))).encode())
	return h.hexdigest()[:16]


# ---------------------------------------------------------------------------------------
# CPU arm.  kind "reference": the UNMODIFIED Python reference (copied by `make -C oracle ref` into oracle/_ref, which
# is git-ignored but travels to the GPU box), driven deterministically, one process per chain -- the
# reference's own parallel model (pymodem.py:140-166).  kind "port": the oracle's C/NumPy restatement, same fan-out.
# ---------------------------------------------------------------------------------------
REF_DIR = os.path.join(REPO, "oracle", "_ref")


def reference_available():
	return os.path.exists(os.path.join(REF_DIR, "modems_codecs", "chain_execute.py"))


def _ref_chain_worker(args):
	ci, audio = args
	import contextlib
	import io
	if REF_DIR not in sys.path:
		sys.path.insert(0, REF_DIR)
	import modems_codecs.chain_builder as cb          # the reference's own modules
	import modems_codecs.chain_execute as ce
	from pymodem_b200 import configs
	line = configs.demod_chains(configs.afsk_1200_ax25_super_opt())[ci]
	t0 = time.perf_counter()
	with contextlib.redirect_stdout(io.StringIO()):
		modem = cb.ModemConfigurator(SAMPLE_RATE, line['modem'])                                   # pymodem.py:79-110
		slicer = cb.SlicerConfigurator(getattr(modem, 'output_sample_rate', SAMPLE_RATE), line['slicer'])
		stream = cb.StreamConfigurator(line['stream'])
		codec = cb.CodecConfigurator(line['codec'], line['object_name'])
		packets = ce.process_chain([line['object_name'], modem, slicer, stream, codec], audio)    # chain_execute.py:6-28
	return time.perf_counter() - t0, len(packets)


def _cpu_chain_worker(args):
	ci, audio = args
	from oracle import oracle as orc
	from pymodem_b200 import configs
	line = configs.demod_chains(configs.afsk_1200_ax25_super_opt())[ci]
	chain = orc.Chain(SAMPLE_RATE, line)
	t0 = time.perf_counter()
	pk = chain.process_chunked(audio, chunk=1 << 20)
	return time.perf_counter() - t0, len(pk)


def cpu_single_core(audio_sample):
	"""All 8 chains sequentially on one core (the scalar port) -> chain-samples/s."""
	t0 = time.perf_counter()
	n = 0
	for ci in range(8):
		_, k = _cpu_chain_worker((ci, audio_sample))
		n += k
	dt = time.perf_counter() - t0
	return 8 * len(audio_sample) / dt, dt, n


def run_reference_arm(args):
	"""--impl reference: the reference's CPU implementation of the path on the box's host cores, one process per
	chain, each step a bounded sample of the workload."""
	import multiprocessing as mp
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return
	from oracle import oracle as orc
	orc.build()
	n_chains = 8
	procs = min(n_chains, os.cpu_count() or 1)
	use_ref = reference_available() and not args.port_only
	ref_s = args.ref_seconds if use_ref else args.cpu_seconds
	audio_port = make_audio(args.cpu_seconds, 0)
	audio = audio_port[:int(ref_s * SAMPLE_RATE)] if use_ref and ref_s <= args.cpu_seconds else make_audio(ref_s, 0)
	worker = _ref_chain_worker if use_ref else _cpu_chain_worker
	ctx = mp.get_context("fork")
	with ctx.Pool(procs) as pool:
		def step(fn, a):
			t0 = time.perf_counter()
			pool.map(fn, [(ci, a) for ci in range(n_chains)])
			return time.perf_counter() - t0
		for _ in range(args.warmup):
			step(worker, audio)
		times = [step(worker, audio) for _ in range(args.steps)]
		# the port beside it (two steps), so that both CPU figures come from the same box and run
		port_times = [step(_cpu_chain_worker, audio_port) for _ in range(2)] if use_ref else None
	total = sum(times)
	value = n_chains * len(audio) * args.steps / total
	kind = "reference" if use_ref else "port"
	what = ("the unmodified Python reference (oracle/_ref: chain_builder + chain_execute.process_chain)" if use_ref
		else "chunked oracle port (numpy.convolve FIRs + C slicer/LFSR/AX.25)")
	sample = f"{ref_s:g} s of the 48 kHz workload x {n_chains} chains per step, {what}, one process per chain"
	line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
		"steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
		"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
		"config": {"workload": "afsk_1200_ax25_super_opt x 48 kHz synthetic AWGN AX.25 audio (bounded sample)",
			"chains": n_chains, "sample_rate": SAMPLE_RATE, "seconds_per_step": ref_s},
		"cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
		"e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0, "host_cpus": os.cpu_count()}
	if port_times:
		line["port"] = {"value": n_chains * len(audio_port) * len(port_times) / sum(port_times), "unit": UNIT, "cores": procs,
			"kind": "port", "sample": f"{args.cpu_seconds:g} s x {n_chains} chains per step, oracle port, one process per chain, 2 steps"}
	print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
def kernel_rooflines(eng, n, n_chains, macs, fp32_peak, peaks, stats, runs=3, tc_macs=0.0):
	"""A timing pass (option kernel_times: a CUDA event before every launch) after the timed region: every kernel with
	at least 2 % of the step gets a roofline line.  The bound named is the one the kernel is built against; `frac` is
	achieved / peak for that bound; for the latency-bound kernels the HBM fraction says how far the bytes are from
	mattering."""
	eng.set_option("kernel_times", 1)
	acc = {}
	order = []
	try:
		for _ in range(runs):
			yield_run = eng._bench_device_step()
			for name, cnt, ms in eng.kernel_times():
				if name not in acc:
					acc[name] = [0, 0.0]
					order.append(name)
				acc[name][0] += cnt
				acc[name][1] += ms
	finally:
		eng.set_option("kernel_times", 0)
	total = sum(v[1] for v in acc.values()) / runs
	hbm = peaks.get("hbm_gbs") or 6650.0
	words = n / 32.0
	st = stats[-1]
	nbits = st["n_stream_bits"]
	flagged = st["guard_flagged"]
	# algorithmic bytes (or flops) per launch group, per step
	sign_mask = n_chains * words * 4.0
	st_seg, st_exact = 24576, 4096          # engine defaults (csrc/engine.cu opt_seg_words, opt_warm_exact_words)
	work = {
		"afsk_front_kernel": ("fp32", 2.0 * macs * n / 1e12, fp32_peak, "TFLOP/s"),
		# tensor-core route: the front kernel stops at the magnitudes (band-pass + sliding correlators + piece split), the
		# low-pass is a tcgen05 GEMM: executed bf16 flops against the measured cuBLAS bf16 peak (burst figure: the kernel is
		# timed alone)
		"afsk_front_kernel (magnitudes)": ("fp32", 2.0 * macs * n / 1e12, fp32_peak, "TFLOP/s"),
		"lpf_tc_kernel": ("tensor", 2.0 * tc_macs * n / 1e12, peaks.get("bf16_tflops") or 1590.0, "TFLOP/s"),
		"guard_fixup_kernel": ("fp64", flagged * 2.0 * (320 * 148 + 4 * 100 * 60 + 100) / 1e12, FP64_NOMINAL_TFLOPS, "TFLOP/s"),
		# the slicer is bound by the issue of its exact step (DSETP, select, DADD, bit test, two selects, DMUL, funnel shift: the
		# float64 and integer pipes of a scheduler take 15 cycles per warp and sample, whatever the occupancy --
		# tools/ubench/slicer_step.cu, profiles/r02af_slicer_step.txt: 2.54e12 thread-steps/s with 24 warps per SM).
		# Executed exact steps of a launch: every segment plus the exact tail of its warm-up, per chain.
		"slicer_segments_kernel": ("fp64+alu issue", n_chains * (n / float(st_seg)) * (st_seg + st_exact) / 1e9, SLICER_STEP_PEAK_G, "G steps/s"),
		"slicer_verify_kernel": ("latency", 0.0, hbm, "GB/s"),
		"slicer_repair_kernel": ("latency", 0.0, hbm, "GB/s"),      # a repair is one thread's chain of up to a segment of exact steps
		"gather_count_kernel": ("hbm", sign_mask / 1e9, hbm, "GB/s"),
		"memset bits": ("hbm", nbits / 8.0 / 1e9, hbm, "GB/s"),
		"gather_write_kernel": ("hbm", (2.0 * sign_mask + nbits / 8.0 + nbits / 8.0 * 4.0) / 1e9, hbm, "GB/s"),
		"lfsr_kernel": ("hbm", 2.0 * nbits / 8.0 / 1e9, hbm, "GB/s"),
		"flag_count_kernel": ("hbm", nbits / 8.0 / 1e9, hbm, "GB/s"),
		"flag_write_kernel": ("hbm", nbits / 8.0 / 1e9, hbm, "GB/s"),
		"ax25_gap_filter_kernel": ("hbm", nbits / 8.0 / 1e9, hbm, "GB/s"),
		"ax25_gap_kernel": ("latency", nbits / 8.0 / 14.0 / 1e9, hbm, "GB/s"),
	}
	out = []
	for name in order:
		cnt, ms = acc[name][0] / runs, acc[name][1] / runs
		row = {"kernel": name, "launches": cnt, "ms": ms, "share": ms / total if total else None}
		if name in work and ms > 0:
			bound, amount, peak, unit = work[name]
			achieved = amount / (ms * 1e-3)
			row.update({"bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
				"frac": achieved / peak if peak else None})
		out.append(row)
	return {"step_ms_under_events": total, "min_share_listed": 0.0,
		"note": "timing pass with an event before every launch; shares, not absolute times, carry over to the timed region; "
			"slicer_segments: achieved = exact float64 steps executed (segment + 4096-sample exact warm-up tail per segment) against "
			"the step rate of tools/ubench/slicer_step.cu (profiles/r02af_slicer_step.txt); the crossing-by-crossing warm-up before "
			"the tail (45056 samples per segment, 0.37 of the kernel's 1.08 ms) is not counted as work; fp64 peak nominal 37.2 TFLOP/s",
		"kernels": out}


def run_b200_arm(args):
	import torch
	import torch.distributed as dist
	from pymodem_b200 import configs
	from pymodem_b200.engine import Engine, engine_for, measure_fp32_peak, pinned_empty
	from pymodem_b200.modems_codecs import chain_builder, chain_execute

	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local = int(os.environ.get("LOCAL_RANK", "0"))
	if not torch.cuda.is_available():
		raise SystemExit("bench.py: no CUDA device -- the demod_chain engine has no CPU fallback")
	torch.cuda.set_device(local)
	from pymodem_b200.sharded import bind_to_gpu_numa_node
	numa = bind_to_gpu_numa_node(local)          # before any pinned allocation
	print(f"[bench rank {rank}] {numa}", file=sys.stderr, flush=True)
	# stdout carries exactly one JSON line: whatever libraries print while the process group and the first communicator
	# come up (NCCL's version banner) goes to stderr
	sys.stdout.flush()
	saved_stdout = os.dup(1)
	os.dup2(2, 1)
	if world > 1:
		dist.init_process_group("nccl", device_id=torch.device("cuda", local))
		dist.barrier()

	def barrier():
		if world > 1:
			dist.barrier()

	strong = args.scaling == "strong"
	lines = configs.afsk_1200_ax25_super_opt()
	stack = [chain_builder.build_chain(SAMPLE_RATE, l) for l in configs.demod_chains(lines)]
	n_chains = len(stack)
	phase_ms = {}
	hour = make_audio(args.seconds, 0)
	n_hour = len(hour)
	opts = dict(kv.split("=") for kv in args.opt)
	eng = Engine(stack, device=local, **opts)
	fp32_peak = measure_fp32_peak(local) if rank == 0 else None
	h2d_ceiling = None

	if world == 1:
		n = n_hour
		host = pinned_empty(n)                # what python -m pymodem_b200 reads the WAV into
		host[:] = hour
		dev_audio = torch.from_numpy(host).cuda(non_blocking=False)
		torch.cuda.synchronize()
		user_eng = engine_for(stack, device=local, **opts)      # the engine process_chains uses (cached per stack)

		def step_device():
			eng.run_device_ptr(dev_audio.data_ptr(), n)
			return eng.stats()
		eng._bench_device_step = step_device

		def step_host():
			# the call a user makes: records on the host, one packet sequence per chain
			per_chain = chain_execute.process_chains(stack, host, device=local, **opts)
			step_host.last = per_chain
			return user_eng.stats()
		n_total = n
		# plain-copy ceiling of this box: the same pinned buffer, one cudaMemcpyAsync, best of 5
		src = torch.from_numpy(host)
		best = None
		for _ in range(5):
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record()
			dev_audio.copy_(src, non_blocking=True)
			e1.record()
			torch.cuda.synchronize()
			ms = e0.elapsed_time(e1)
			best = ms if best is None else min(best, ms)
		h2d_ceiling = {"ms": best, "gbs": 2.0 * n / best / 1e6, "bytes": 2 * n, "how": "one cudaMemcpyAsync of the pinned recording, best of 5"}
	else:
		# ONE recording sharded on the sample axis: rank r gets its slice + FIR/warm-up history + a few forward symbols
		# (pymodem_b200/sharded.py).  weak: world x seconds (the synthetic hour repeated); strong: the hour itself.
		from pymodem_b200.sharded import LinkedRun, TorchExchange, choose_segment_len, plan_shards
		n_total = n_hour if strong else n_hour * world
		# shorter slicer segments when a rank's shard is small (strong scaling): see sharded.choose_segment_len
		seg_len = int(float(opts["segment_len"])) if "segment_len" in opts else choose_segment_len(n_total // world, n_chains)
		eng.set_option("segment_len", seg_len)
		plans = plan_shards(n_total, world, segment_len=seg_len, trim_max=305, samples_per_symbol=40.0, tail_bits=16384)
		plan = plans[rank]
		max_local = max(p['audio_end'] - p['audio_begin'] for p in plans)      # the link layout must be the same on every rank
		idx = np.arange(plan['audio_begin'], plan['audio_end'], dtype=np.int64) % n_hour
		n = len(idx)
		host = pinned_empty(n)
		host[:] = hour[idx]
		del idx
		dev_audio = torch.from_numpy(host).cuda(non_blocking=False)
		torch.cuda.synchronize()
		# the ranks trade CUDA IPC handles once (NCCL all-gather); after that the hand-off, the bit tails and the
		# packet records travel over NVLink peer memory inside each rank's own kernel stream (csrc/link.cu).
		# NCCL is used again only if a speculated slicer start state does not verify (repair protocol) or a shard
		# cannot finish its decode from what it holds (bitstream recovery).
		ex = TorchExchange(torch.device("cuda", local))
		linked = LinkedRun(eng, rank, world, max_local, ex, ex.var, tail_bits=16384)

		def step_device():
			linked.run(plan, dev_audio.data_ptr(), n, on_device=True, timing=phase_ms, fetch=False)
			return eng.stats()

		def step_host():
			res = linked.run(plan, host.ctypes.data, n, on_device=False, timing=phase_ms, fetch=True)
			step_host.last = eng.packets(*res)
			step_host.raw = res
			return eng.stats()

	def timed(step, k):
		phase_ms.clear()
		barrier()
		torch.cuda.synchronize()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		t0 = time.perf_counter()
		stats = [step() for _ in range(k)]          # every step ends with results on the host (stream sync)
		e1.record()
		torch.cuda.synchronize()
		wall = time.perf_counter() - t0
		ms = max(e0.elapsed_time(e1), 0.0)
		t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
		if world > 1:
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
		barrier()
		return float(t[0]), float(t[1]), stats

	for _ in range(args.warmup):
		step_device()
	sys.stdout.flush()
	os.dup2(saved_stdout, 1)
	os.close(saved_stdout)
	sampler = ClockSampler(local)
	if rank == 0:
		sampler.start()
	ev_ms, wall_ms, stats = timed(step_device, args.steps)
	shard_phase = {k: v / args.steps for k, v in phase_ms.items()}
	# e2e: the same metric through the user-facing call with host buffers
	for _ in range(min(args.warmup, 3)):
		step_host()
	e2e_ev_ms, e2e_wall_ms, e2e_stats = timed(step_host, args.steps)
	e2e_phase = {k: v / args.steps for k, v in phase_ms.items()}

	# Variants of the host-buffer path (N=1), each a complete pass per step:
	#  pageable : the caller's buffer is ordinary NumPy memory -> staged through the engine's pinned ring by host threads
	#  pipelined: two recordings in flight (two engines, two host threads): the H2D copy of one overlaps the slicer /
	#             bit-level tail of the other, so the PCIe link never idles
	pageable = pipelined = None
	if world == 1:
		pg = hour.copy()
		def step_pageable():
			chain_execute.process_chains(stack, pg, device=local, **opts)
			return user_eng.stats()
		for _ in range(2):
			step_pageable()
		k = max(args.steps // 4, 2)
		p_ev, p_wall, _ = timed(step_pageable, k)
		pageable = {"value": n_chains * n / (max(p_ev, p_wall) / k * 1e-3), "unit": UNIT, "ms_per_step": max(p_ev, p_wall) / k,
			"note": "process_chains on a pageable NumPy array: chunks staged through the engine's pinned ring (memcpy on host threads)"}
		del pg
		if args.in_flight > 1:
			engs = [user_eng] + [Engine(stack, device=local, **opts) for _ in range(args.in_flight - 1)]
			for e2 in engs[1:]:
				for _ in range(3):
					e2.run(host)
			per = max(args.steps // len(engs), 2)
			counts = []

			def worker(e2):
				for _ in range(per):
					counts.append(sum(len(p) for p in e2.run(host)))
			threads = [threading.Thread(target=worker, args=(e2,)) for e2 in engs]
			torch.cuda.synchronize()
			t0 = time.perf_counter()
			for th in threads:
				th.start()
			for th in threads:
				th.join()
			torch.cuda.synchronize()
			dt = time.perf_counter() - t0
			pipelined = {"in_flight": len(engs), "steps": per * len(engs), "ms_per_step": 1e3 * dt / (per * len(engs)),
				"value": n_chains * n * per * len(engs) / dt, "unit": UNIT,
				"note": "Engine.run(host buffer) from two host threads, two recordings in flight on one GPU: every step is a "
					"complete pass (H2D, kernels, records D2H, packet sequences); throughput, not latency"}
			for e2 in engs[1:]:
				e2.close()
	clocks = sampler.stop() if rank == 0 else None

	# ---- parity of what the timed steps produced (outside the timed region) ----
	parity = None
	if world == 1:
		per_chain = step_host.last
		got = packets_digest(per_chain)
		default_hour = args.seconds == 3600.0 and not args.opt
		parity = {"n_records": sum(len(p) for p in per_chain), "digest": got,
			"oracle_digest": ORACLE_HOUR_DIGEST if default_hour else None,
			"match": (got == ORACLE_HOUR_DIGEST) if default_hour else None,
			"what": "packet set (chain, streamaddress, payload bytes, BytesCorrected) of the last e2e step against the CPU oracle's "
				"for the same hour (profiles/r01g_verify_hour.txt, tools/verify_hour.py)"}
	else:
		recs, arena = step_host.raw
		mine = records_digest(recs, arena)
		ref = None
		if rank == 0:
			# the same recording, unsharded, on this rank's GPU alone
			whole = hour if strong else np.tile(hour, world)
			solo = Engine(stack, device=local, **opts)
			try:
				ref = records_digest(*solo.run_raw(whole))
			finally:
				solo.close()
			del whole
		digests = [None] * world
		dist.all_gather_object(digests, mine)
		parity = {"n_records": int(len(recs)), "digest_per_rank": digests, "unsharded_digest": ref,
			"match": bool(ref is not None and all(d == ref for d in digests)),
			"what": "sha256 of the merged packet records every rank ends a step with, against an unsharded single-GPU run of the "
				"same recording on rank 0"}

	n_packets = stats[-1]["n_packets"]
	link_fallbacks = linked.fallbacks if world > 1 else 0
	launches = sum(s["kernel_launches"] for s in stats)
	front_ms = statistics.mean(s["front_ms"] for s in stats)
	front_launches = stats[-1]["front_launches"]
	total_units = n_chains * n_total
	ms_per_step = max(ev_ms, wall_ms) / args.steps
	value = total_units / (ms_per_step * 1e-3)
	e2e_ms = max(e2e_ev_ms, e2e_wall_ms) / args.steps
	e2e_value = total_units / (e2e_ms * 1e-3)

	if rank != 0:
		eng.close()
		if world > 1:
			dist.destroy_process_group()
		return

	# roofline of the dominant kernel (afsk_front_kernel: FP32 FFMA bound)
	macs = eng.front_macs_per_sample()
	executed_flop = 2.0 * macs * n
	t_front = front_ms * 1e-3 / max(front_launches, 1)
	achieved = executed_flop / max(front_launches, 1) / t_front / 1e12
	peaks, peaks_src = measured_peaks()
	algo_bytes = 2.0 * n + n_chains * n / 8.0        # int16 audio in, 1 sign bit per chain-sample out
	def committed_traffic(kernel):
		"""DRAM bytes of a kernel from the committed ncu --set full capture (profiles/kernel_traffic.json), scaled to this
		launch's samples; None when the capture does not have it."""
		try:
			with open(os.path.join(REPO, "profiles", "kernel_traffic.json")) as f:
				tr = json.load(f)
			k = tr["kernels"][kernel]
			return (k["dram_bytes_read"] + k["dram_bytes_write"]) * (n / tr["samples_per_launch"])
		except (OSError, ValueError, KeyError):
			return None
	traffic = committed_traffic("afsk_front_kernel")
	tc_macs = eng.front_tensor_macs_per_sample()
	roofline = {"kernel": "afsk_front_kernel" + (" (magnitudes) + lpf_tc_kernel" if tc_macs else ""), "bound": "fp32", "achieved": achieved, "peak": fp32_peak,
		"unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None, "traffic": traffic,
		"peak_source": "pm_measure_fp32_peak: register-resident FFMA loop timed in this run "
			f"(nominal {FP32_NOMINAL_TFLOPS:.1f} at max clocks; MEASURED_PEAKS.json has no FP32 entry)",
		"executed_mac_per_sample": macs,
		"reference_equiv_tflops": REF_FLOP_PER_CHAIN_SAMPLE * n_chains * n / t_front / max(front_launches, 1) / 1e12,
		"kernel_ms": front_ms / max(front_launches, 1), "launches_per_step": front_launches,
		"hbm": {"algorithmic_bytes": algo_bytes, "achieved_gbs": algo_bytes / t_front / 1e9,
			"peak_gbs": peaks.get("hbm_gbs"), "peak_source": peaks_src}}
	stage_ms = {k: statistics.mean(s[k] for s in stats) for k in ("total_ms", "front_ms", "fixup_ms", "slicer_ms", "bits_ms", "d2h_ms")}
	roofline_all = None
	if world == 1:
		try:
			roofline_all = kernel_rooflines(eng, n, n_chains, macs, fp32_peak, peaks, stats, tc_macs=tc_macs)
			# the headline roofline is that of the kernel with the largest share of the step, timed in that pass (with the
			# tensor-core low-pass the front end is two kernels and front_ms above covers both)
			top = max((k for k in roofline_all["kernels"] if k.get("frac")), key=lambda k: k["ms"])
			roofline.update({"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"],
				"unit": top["unit"], "frac": top["frac"], "kernel_ms": top["ms"], "launches_per_step": top["launches"],
				"share_of_step": top["share"], "traffic": committed_traffic(top["kernel"])})
			if top["kernel"] == "afsk_front_kernel (magnitudes)":
				# what the kernel has to move: int16 audio in, four tone magnitudes out as three bf16 pieces each
				roofline["hbm"]["algorithmic_bytes"] = 2.0 * n + 4 * 3 * 2.0 * n
				roofline["hbm"]["achieved_gbs"] = roofline["hbm"]["algorithmic_bytes"] / (top["ms"] * 1e-3) / 1e9
			if top["bound"] != "fp32":
				roofline["peak_source"] = peaks_src
			if tc_macs:
				tck = next((k for k in roofline_all["kernels"] if k["kernel"] == "lpf_tc_kernel"), None)
				if tck:
					roofline["tensor"] = {"kernel": "lpf_tc_kernel", "bound": "tensor", "achieved": tck["achieved"], "peak": tck["peak"],
						"unit": "TFLOP/s", "frac": tck["frac"], "kernel_ms": tck["ms"], "peak_sustained": peaks.get("bf16_tflops_sustained"),
						"executed_bf16_mac_per_sample": tc_macs,
						"useful_mac_per_sample": eng._lib.pm_engine_front_lpf_macs_per_sample(eng._h),
						"note": "three bf16 pieces per operand, six piece products, K padded to the Toeplitz band: executed flops, "
							"not useful ones; peak = cuBLAS bf16 burst figure of MEASURED_PEAKS.json"}
		except Exception as exc:                 # the timing pass is informational: never lose the line over it
			roofline_all = {"error": f"{type(exc).__name__}: {exc}"}

	# CPU baseline on a bounded sample of the same workload (rank 0, N=1 only)
	cpu = None
	if world == 1 and not args.no_cpu:
		from oracle import oracle as orc
		orc.build()
		sample = hour[: int(args.cpu_seconds * SAMPLE_RATE)]
		v, dt, _ = cpu_single_core(sample)
		cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
			"sample": f"first {len(sample) / SAMPLE_RATE:g} s of this workload x {n_chains} chains, oracle port "
				f"(numpy.convolve + C slicer/LFSR/AX.25) on one core, {dt:.1f} s"}

	e2e = {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
		"h2d_bytes_per_step": e2e_stats[-1]["h2d_bytes"], "d2h_bytes_per_step": e2e_stats[-1]["d2h_bytes"],
		"bytes_scope": "rank 0 (every rank copies its own shard of the recording plus the slicer's warm-up history)",
		"call": ("chain_execute.process_chains(demod_stack, audio in pinned host memory) -> per-chain packet sequences" if world == 1 else
			"LinkedRun.run(shard in pinned host memory) -> merged records of all ranks -> per-chain packet sequences"),
		"frac_of_value": e2e_value / value, "pageable": pageable, "pipelined": pipelined,
		"shard_phase_ms_rank0": e2e_phase}
	if h2d_ceiling:
		e2e["h2d_ceiling"] = h2d_ceiling
		e2e["h2d_gbs"] = e2e_stats[-1]["h2d_bytes"] / (e2e_ms * 1e-3) / 1e9
		e2e["pcie_frac"] = h2d_ceiling["ms"] / e2e_ms       # share of the step the plain copy alone would take
		if pipelined:
			pipelined["pcie_frac"] = h2d_ceiling["ms"] / pipelined["ms_per_step"]

	line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
		"ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
		"dtype": "f32", "data": "synthetic",
		"config": {"workload": f"afsk_1200_ax25_super_opt ({n_chains} chains) x {args.seconds:g} s of 48 kHz synthetic "
			f"AWGN AX.25 audio {'in total' if strong else 'per GPU'}", "chains": n_chains, "sample_rate": SAMPLE_RATE,
			"samples_per_gpu": n_total // world if strong else n_hour, "samples_total": n_total,
			"l2": "inputs larger than L2 (345.6 MB int16 audio per GPU per step)" if n * 2 > 126e6 else "input fits L2",
			"recording": ("one recording of `seconds`, sharded on the sample axis" if strong else
				"one recording of n_gpus x seconds (the synthetic hour repeated), sharded on the sample axis"),
			"parallelism": f"chains x audio segments on {world} GPU(s); one shard of the recording per rank; slicer "
				"hand-off, bit tails and packet records exchanged by the GPUs over NVLink peer memory (csrc/link.cu), "
				"every rank ends the step with the merged records of all ranks on its host"},
		"e2e": e2e, "gpu_launches": launches, "roofline": roofline, "roofline_all": roofline_all, "cpu_baseline": cpu,
		"clocks": clocks, "parity": parity,
		"stage_ms": stage_ms, "shard_phase_ms_rank0": shard_phase, "link_fallbacks": link_fallbacks,
		"link_recoveries": linked.recoveries if world > 1 else 0, "packets_per_step": n_packets,
		"slicer": {"segments": stats[-1]["slicer_segments"], "repairs": stats[-1]["slicer_repairs"],
			"guard_flagged": stats[-1]["guard_flagged"]},
		"timing": {"cuda_event_ms": ev_ms, "wall_ms": wall_ms, "e2e_cuda_event_ms": e2e_ev_ms, "e2e_wall_ms": e2e_wall_ms},
		"host_cpus": os.cpu_count(), "numa": numa}
	print(json.dumps(line), flush=True)
	eng.close()
	if world > 1:
		dist.destroy_process_group()


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=20)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
	ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
		help="weak: `seconds` of audio per GPU; strong: `seconds` in total, split over the GPUs")
	ap.add_argument("--config", default="super_opt", help="super_opt (the headline) or one of the other shipped configs: see bench_configs.py")
	ap.add_argument("--seconds", type=float, default=3600.0, help="audio per GPU per step (in total with --scaling strong)")
	ap.add_argument("--cpu-seconds", type=float, default=60.0, help="bounded sample for the CPU port")
	ap.add_argument("--ref-seconds", type=float, default=20.0, help="bounded sample per step for the Python reference (--impl reference)")
	ap.add_argument("--port-only", action="store_true", help="--impl reference: time the oracle port even when oracle/_ref exists")
	ap.add_argument("--no-cpu", action="store_true")
	ap.add_argument("--in-flight", type=int, default=2, help="recordings in flight for e2e.pipelined (N=1 only; 1 = skip)")
	ap.add_argument("--opt", action="append", default=[], help="engine option key=value (pm_engine_set_option)")
	ap.add_argument("--batch", type=int, default=0, help="--config lines: recordings per engine call (0 = the config's default)")
	args = ap.parse_args()
	args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
	if args.config != "super_opt":
		import bench_configs
		return bench_configs.main(args)
	if args.impl == "reference":
		run_reference_arm(args)
	else:
		run_b200_arm(args)


if __name__ == "__main__":
	main()
