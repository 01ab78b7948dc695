"""One recording sharded on the sample axis over several GPUs (one engine = one rank).

The reference runs every chain over the whole file (pymodem.py:140-166).  Here chains stay
together (they share the front end) and the *audio* is split: rank r owns the soft samples
[B_r, B_{r+1}) and receives, besides its slice, the FIR history + slicer warm-up before it and
a few symbols after it.  Three tiny exchanges make the result bit-identical to the unsharded
run (all of them all-gathers of a few KB; NCCL on GPUs, gloo in the CPU tests):

  1. slicer hand-off  -- every rank speculates its start state from the warm-up; the previous
     rank's true end state is compared bit for bit and the shard is repaired when it differs
     (repeated until no end state changes: at most `world` rounds, normally one);
  2. symbol counts    -- stream bytes are cut every 8 bits counted from sample 0
     (slicer.py:92-97), so a rank needs the number of symbols of all earlier ranks;
  3. bit tails        -- the last `tail_bits` bits of every rank go to the next one, so a frame
     that straddles a boundary is decoded by the rank that holds its closing flag.

Finally the packet records of all ranks are gathered (rank order == stream order) for the
existing streamaddress-based Correlate on the host (packet_meta.py:230-271).
"""
import ctypes
import math
import struct

import numpy as np

from . import _lib
from .engine import REC_DTYPE


# the engine's slicer geometry defaults (csrc/engine.cu opt_seg_words / opt_warm_words): shard boundaries are segment aligned
DEFAULT_SEGMENT_LEN = 24576
DEFAULT_WARMUP_LEN = 49152


def choose_segment_len(n_local, n_chains, target_threads=49152, lo=8192, hi=DEFAULT_SEGMENT_LEN):
	"""Slicer segment length for a shard of n_local samples: the slicer kernel is one thread per (chain, segment) and
	bound by the latency of a thread's dependent chain (warm-up + segment), so a small shard -- one hour split over 8
	GPUs -- wants shorter segments than the 24576 samples that are best when a GPU has the whole hour: enough threads
	to fill the machine, a shorter chain per thread.  Results do not depend on it (every hand-off is verified).  All
	ranks of a run must use the same value (shard boundaries are segment aligned): compute it from the largest shard.
	Measured on a 450 s shard (tools/slicer_sweep_short.py, profiles/r02_slicer_sweep_short.txt): 24576 -> 1.36 ms, 8192 -> 1.10 ms,
	4096 -> 1.02 ms of slicer but more segments to gather; shorter warm-ups than 49152 samples cost more in repairs than they
	save (a warm-up has to see ~90 zero crossings before the clock is bit-identical)."""
	want = max(1, n_local * max(n_chains, 1) // target_threads)
	seg = max(lo, min(hi, want))
	return max(1024, seg // 1024 * 1024)


def plan_shards(n_samples, world, segment_len=DEFAULT_SEGMENT_LEN, warm_len=DEFAULT_WARMUP_LEN, trim_max=305, samples_per_symbol=40.0,
		tail_bits=16384, pre_segments=None, pre_samples=4 * DEFAULT_SEGMENT_LEN):
	"""Split n_samples over `world` ranks.  Returns one dict per rank:
	audio_begin/audio_end (the slice of the recording the rank needs) + the pm_shard_plan fields.

	pre_segments: whole segments of slicer history every later rank runs (and verifies) before its own range, so that
	its speculated state at own_begin rests on (pre_segments + 1) warm-ups instead of one.  A single
	warm-up fails to become bit-identical in ~1 % of the cases (profiles/r01_slicer_sweep.txt); inside one GPU that
	costs a cheap re-run of the segment, at a shard boundary it would cost the hand-off's fast path.  The history is
	0.1 % more front-end work per rank."""
	if world < 1:
		raise ValueError("world must be >= 1")
	if pre_segments is None:
		# the verified history before a shard is a number of SAMPLES (98304 by default: four default segments), whatever
		# the segment length: with short segments and the same count the hand-off would rest on too little and fall back
		pre_segments = int(math.ceil(pre_samples / segment_len))
	per = int(math.ceil(n_samples / world / segment_len)) * segment_len
	if world > 1 and per * (world - 1) >= n_samples - trim_max:
		raise ValueError(f"recording too short ({n_samples} samples) for {world} shards of segment_len {segment_len}")
	back = int(math.ceil(warm_len / segment_len)) * segment_len
	fwd = int(math.ceil(16 * samples_per_symbol / 32.0)) * 32 + 64
	plans = []
	for r in range(world):
		b0 = r * per
		pre = 0 if r == 0 else max(0, min(pre_segments, (b0 - back) // segment_len))
		base = max(0, b0 - back - pre * segment_len)
		last = r == world - 1
		a1 = n_samples if last else min(n_samples, b0 + per + fwd + trim_max)
		plans.append(dict(rank=r, audio_begin=base, audio_end=a1, sample_base=base, own_begin=b0 - base,
			own_len=per, first=(r == 0), last=last, tail_bits=tail_bits if world > 1 else 0, pre_segments=pre))
	return plans


def _states_to_bytes(states):
	return bytes(ctypes.string_at(ctypes.addressof(states), ctypes.sizeof(states)))


def _states_from_bytes(blob, n_chains):
	arr = (_lib.ShardState * n_chains)()
	ctypes.memmove(ctypes.addressof(arr), blob, ctypes.sizeof(arr))
	return arr


_STATE_DTYPE = np.dtype([('start_clock', '<u8'), ('end_clock', '<u8'), ('start_last', '<u4'), ('start_last_q', '<u4'),
	('end_last', '<u4'), ('end_last_q', '<u4'), ('n_symbols', '<i8')])      # pm_shard_state, clocks as bit patterns
assert _STATE_DTYPE.itemsize == ctypes.sizeof(_lib.ShardState)


def handoffs_verified(all_blobs):
	"""True when every rank's speculated start state is bit for bit the previous rank's end state, i.e. when no
	rank has anything to repair.  Every rank evaluates this on the same gathered data, so all of them take the
	same branch without another collective."""
	st = [np.frombuffer(b, dtype=_STATE_DTYPE) for b in all_blobs]
	for prev, cur in zip(st, st[1:]):
		if (not np.array_equal(prev['end_clock'], cur['start_clock']) or not np.array_equal(prev['end_last'], cur['start_last'])
				or not np.array_equal(prev['end_last_q'], cur['start_last_q'])):
			return False
	return True


class ShardWorker:
	"""The per-rank side of the protocol; `engine` is a pymodem_b200.engine.Engine (or, in the
	CPU tests, an object with the same shard_* / fetch methods)."""

	def __init__(self, engine, plan, audio_ptr, n_local, on_device=False):
		self.engine, self.plan, self.audio_ptr, self.n_local, self.on_device = engine, plan, audio_ptr, n_local, on_device
		self.rank = plan['rank']
		self.n_chains = engine.n_chains
		self.rounds = 0

	def begin(self):
		self.states = self.engine.shard_begin(self.audio_ptr, self.n_local, self.plan, self.on_device)
		return _states_to_bytes(self.states)

	def resume(self):
		"""The states of a run that was begun through the shard link (Engine.run_linked_begin) and did not verify."""
		self.states = self.engine.shard_states()
		return _states_to_bytes(self.states)

	def handoff(self, all_blobs):
		prev = _states_from_bytes(all_blobs[self.rank - 1], self.n_chains) if self.rank > 0 else None
		self.states, changed = self.engine.shard_handoff(prev)
		self.rounds += 1
		return _states_to_bytes(self.states), changed

	def gather(self, all_blobs):
		before = [0] * self.n_chains
		for q in range(self.rank):
			st = _states_from_bytes(all_blobs[q], self.n_chains)
			for c in range(self.n_chains):
				before[c] += int(st[c].n_symbols)
		tail = self.engine.shard_gather(before)
		return np.ascontiguousarray(tail, dtype=np.uint32).tobytes()

	def finish(self, all_tails, il2p_prev=None, il2p=False):
		"""il2p: the engine has IL2P chains -- il2p_prev is the previous rank's decoder-state blob (None on rank 0) and
		self.il2p_blob becomes this rank's.  A shard that cannot finish from what it holds (a frame reaching back
		past the hand-off tail, the max_packet_length overflow of ax25.py:46-51, an IL2P frame longer than the tail)
		answers RECOVER: the ranks then decode the gathered bitstream together (recover_from_bitstream)."""
		from .engine import EngineError
		tail_in = None
		if self.rank > 0 and self.plan['tail_bits'] > 0:
			tail_in = np.frombuffer(all_tails[self.rank - 1], dtype=np.uint32).reshape(self.n_chains, -1)
		try:
			if il2p:
				self.il2p_blob = bytes(ctypes.sizeof(_lib.Il2pState) * self.n_chains)
				prev = None
				if il2p_prev is not None:
					prev = (_lib.Il2pState * self.n_chains)()
					ctypes.memmove(ctypes.addressof(prev), il2p_prev, ctypes.sizeof(prev))
				out = self.engine.shard_finish_il2p(tail_in, prev)
				self.il2p_blob = bytes(ctypes.string_at(ctypes.addressof(out), ctypes.sizeof(out)))
			else:
				self.engine.shard_finish(tail_in)
		except EngineError as exc:
			if f"error {_lib.PM_ERR_STATE}:" not in str(exc):
				raise
			return RECOVER
		recs, arena = self.engine.fetch()
		return struct.pack("<qq", len(recs), len(arena)) + recs.tobytes() + arena.tobytes()


RECOVER = b"RECOVER!" * 2       # in place of a rank's result blob (16 bytes: never a valid header + records)


def export_blob(engine):
	"""What one rank contributes to the recovery: for every chain the local stream bits, byte addresses and their
	placement (Engine.shard_export)."""
	parts = [struct.pack("<q", engine.n_chains)]
	for c in range(engine.n_chains):
		bits, addr, info = engine.shard_export(c)
		parts.append(struct.pack("<6q", *info, len(bits), len(addr)) + bits.tobytes() + addr.tobytes())
	return b"".join(parts)


def assemble_streams(blobs):
	"""Export blobs of all ranks in rank order -> per chain (bytes uint8[n], addresses int64[n]): the AddressedData
	stream the unsharded slicer produces (slicer.py:92-97): every rank's own bits back to back, cut into bytes every 8
	bits counted from the start of the recording, each byte carrying the address of the sample that completed it
	(held by the rank that owns the byte's last bit); a trailing partial byte is dropped."""
	n_chains = struct.unpack_from("<q", blobs[0], 0)[0]
	offs = [8] * len(blobs)
	out = []
	for _c in range(n_chains):
		pieces, placed, pos = [], [], 0
		for r, blob in enumerate(blobs):
			nbits, bit_off, n_own, base, nw, na = struct.unpack_from("<6q", blob, offs[r])
			o = offs[r] + 48
			words = np.frombuffer(blob, dtype=np.uint32, count=nw, offset=o)
			addr = np.frombuffer(blob, dtype=np.uint32, count=na, offset=o + 4 * nw)
			offs[r] = o + 4 * nw + 4 * na
			local = np.unpackbits(words.view(np.uint8), bitorder='little')
			pieces.append(local[bit_off:bit_off + n_own])
			a0 = pos - bit_off                      # global index of local bit 0: a multiple of 8 (byte cuts are global)
			assert a0 % 8 == 0 and a0 >= 0, "shard placement is not byte aligned"
			j_lo, j_hi = bit_off // 8, (bit_off + n_own) // 8      # local bytes whose last bit is an own bit
			placed.append((a0 // 8 + j_lo, addr[j_lo:j_hi].astype(np.int64) + base))
			pos += n_own
		nbytes = pos // 8
		bits = np.concatenate(pieces)[:8 * nbytes] if pieces else np.zeros(0, dtype=np.uint8)
		data = np.packbits(bits, bitorder='big')
		addresses = np.zeros(nbytes, dtype=np.int64)
		for first, vals in placed:
			vals = vals[:max(0, nbytes - first)]
			addresses[first:first + len(vals)] = vals
		out.append((data, addresses))
	return out


def decode_streams(engine, streams):
	"""Per-chain AddressedData streams -> (records, arena) ordered like an unsharded run, through the engine's own
	descrambler and codec kernels (Engine.unscramble_stream / decode_stream)."""
	recs_all, arenas, base = [], [], 0
	for c, (data, addresses) in enumerate(streams):
		d1, a1 = engine.unscramble_stream(c, data, addresses)
		recs, arena = engine.decode_stream(c, d1, a1)
		recs = recs.copy()
		recs['offset'] += base
		base += len(arena)
		recs_all.append(recs)
		arenas.append(arena.copy())
	recs = np.concatenate(recs_all) if recs_all else np.zeros(0, dtype=REC_DTYPE)
	arena = np.concatenate(arenas) if arenas else np.zeros(0, dtype=np.uint8)
	return recs, arena


def recover_from_bitstream(engines, exchange_var):
	"""The always-exact way out of a sharded run: every rank exports its part of the sliced stream, the parts are
	gathered (a few MB per chain-hour), and the complete stream of every chain is descrambled and decoded from the
	start -- the same bit-level work an unsharded run does, including the sequential AX.25 replay.  Every rank ends
	with the same (records, arena)."""
	blobs = exchange_var([export_blob(e) for e in engines])
	return decode_streams(engines[0], assemble_streams(blobs))


class Recovered:
	"""Result of a run that went through recover_from_bitstream, shaped like Gathered."""

	def __init__(self, recs, arena):
		self.recs, self.arena = recs, arena

	def n_packets(self):
		return len(self.recs)

	def merge(self):
		return self.recs, self.arena


def _unpack_result(blob):
	"""-> (records as uint64[n, 5] rows -- 40-byte pm_packet_rec, column 0 = chain | len << 32, column 1 = arena
	offset -- and the arena bytes); views into the blob, nothing is copied."""
	n, nb = struct.unpack_from("<qq", blob, 0)
	off = 16
	rows = np.frombuffer(blob, dtype=np.uint64, count=n * 5, offset=off).reshape(n, 5)
	arena = np.frombuffer(blob, dtype=np.uint8, count=nb, offset=off + n * REC_DTYPE.itemsize)
	return rows, arena


_merge_buf = {}


def _scratch(name, n, dtype):
	"""Reused output buffers: a fresh multi-megabyte numpy array costs more in page faults than the copy into it."""
	buf = _merge_buf.get(name)
	if buf is None or len(buf) < n:
		buf = _merge_buf[name] = np.empty(max(n, 1) * 5 // 4 + 64, dtype=dtype)
	return buf[:n]


def merge_results(blobs, n_chains=None):
	"""Per-rank (records, arena) blobs in rank order -> one (records, arena) ordered by (chain, stream position),
	offsets rebased into the merged arena.  Every rank's records are already sorted by (chain, position) and ranks
	are in stream order, so the merge is a block interleave: chain c of rank 0, chain c of rank 1, ...  The records
	are handled as plain uint64 rows (fancy indexing a structured dtype with a sub-array field costs ~100 ns each).
	The returned arrays are views of module-level scratch buffers: valid until the next merge."""
	parts = [_unpack_result(b) for b in blobs]
	total = sum(len(p[0]) for p in parts)
	rows = _scratch('rows', total * 5, np.uint64).reshape(total, 5)
	chains = [r.view(np.uint32)[:, 0] for r, _ in parts]        # low half of column 0, strided view
	if n_chains is None:
		n_chains = 1 + max([int(c[-1]) for c in chains if len(c)], default=0)
	keys = np.arange(n_chains + 1, dtype=np.uint32)
	bounds = [np.searchsorted(c, keys) for c in chains]
	bases, base = [], 0
	for _, arena in parts:
		bases.append(base)
		base += len(arena)
	pos = 0
	counts, shifts = [], []
	for c in range(n_chains):
		for (r, _), b, ab in zip(parts, bounds, bases):
			lo, hi = int(b[c]), int(b[c + 1])
			if hi > lo:
				rows[pos:pos + hi - lo] = r[lo:hi]
				pos += hi - lo
				counts.append(hi - lo)
				shifts.append(ab)
	if counts:
		rows[:, 1] += np.repeat(np.array(shifts, dtype=np.uint64), counts)
	arena = _scratch('arena', base, np.uint8)
	if parts:
		np.concatenate([p[1] for p in parts], out=arena)
	return rows.view(REC_DTYPE).reshape(-1), arena


class Gathered:
	"""The packet records of all ranks as they arrived (one blob per rank, rank order = stream order), on the host
	of every rank.  merge() turns them into one (records, arena) pair ordered like a single engine's output; it is
	host bookkeeping on the consumer's side, like Engine.fetch()/packets() after an unsharded run."""

	def __init__(self, blobs):
		self.blobs = blobs

	def n_packets(self):
		return sum(struct.unpack_from("<q", b, 0)[0] for b in self.blobs)

	def merge(self):
		return merge_results(self.blobs)


def run_protocol(workers, exchange, exchange_var=None, timing=None, merge=True, resume=False):
	"""Drive the shard protocol.  `workers` are the ShardWorkers living in this process (one per
	rank under torch.distributed; all of them when several shards are emulated in one process).
	exchange(list of equally long local blobs) -> the blobs of ALL ranks in rank order.
	Three collectives per run when every speculated hand-off verifies (the normal case): slicer states + symbol
	counts, bit tails, packet records."""
	import time
	exchange_var = exchange_var or exchange
	t = [time.perf_counter()]

	def lap(name):
		t.append(time.perf_counter())
		if timing is not None:
			timing[name] = timing.get(name, 0.0) + (t[-1] - t[-2]) * 1e3
	blobs = [w.resume() if resume else w.begin() for w in workers]
	lap("begin")
	blobs = exchange(blobs)
	lap("exchange")
	while True:
		outs = [w.handoff(blobs) for w in workers]
		lap("handoff")
		if handoffs_verified(blobs):          # nobody had to repair: the states every rank holds are final
			break
		blobs = exchange([o[0] for o in outs])
		lap("exchange")
	tails = [w.gather(blobs) for w in workers]
	lap("gather")
	tails = exchange(tails)
	lap("exchange")
	if not any(getattr(w.engine, 'has_il2p', False) for w in workers):
		results = [w.finish(tails) for w in workers]
	else:
		# IL2P keeps decoder state across a boundary (searching / inside a frame, leaked correction counts): the shards
		# finish one after the other, each handing its state to the next (one small all-gather per rank)
		world = len(tails)
		size = ctypes.sizeof(_lib.Il2pState) * workers[0].n_chains
		state, done = None, {}
		for step in range(world):
			blob = bytes(size)
			for w in workers:
				if w.rank == step:
					done[w.rank] = w.finish(tails, il2p_prev=state, il2p=True)
					blob = w.il2p_blob
			state = exchange([blob] if len(workers) == 1 else [w.il2p_blob if w.rank == step else bytes(size) for w in workers])[step]
		results = [done[w.rank] for w in workers]
	lap("finish")
	results = exchange_var(results)
	lap("exchange")
	if any(r == RECOVER for r in results):
		# every rank sees the same blobs, so every rank takes this branch
		recs, arena = recover_from_bitstream([w.engine for w in workers], exchange_var)
		lap("recover")
		if timing is not None:
			timing['recovered'] = timing.get('recovered', 0) + 1
		return (recs, arena) if merge else Recovered(recs, arena)
	if not merge:
		return Gathered(results)
	merged = merge_results(results)
	lap("merge")
	return merged


def local_exchange(blobs):
	"""All shards live in this process (tests, single-GPU emulation)."""
	return list(blobs)


class TorchExchange:
	"""All-gather of byte blobs over torch.distributed (NCCL through pinned/CUDA staging tensors,
	gloo on CPU tensors).  Staging buffers are cached per blob size."""

	def __init__(self, device):
		import torch
		import torch.distributed as dist
		self.torch, self.dist, self.device = torch, dist, torch.device(device)
		self.world = dist.get_world_size()
		self.cuda = self.device.type == "cuda"
		self._bufs = {}
		self._cap = 1 << 16          # common blob capacity of var(), a multiple of 64 KiB

	def _staging(self, n):
		torch = self.torch
		if n not in self._bufs:
			h_in = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=self.cuda)
			h_out = torch.empty(max(n, 1) * self.world, dtype=torch.uint8, pin_memory=self.cuda)
			d_in = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device) if self.cuda else h_in
			d_out = torch.empty(max(n, 1) * self.world, dtype=torch.uint8, device=self.device) if self.cuda else h_out
			self._bufs[n] = (h_in, h_out, d_in, d_out)
		return self._bufs[n]

	def __call__(self, blobs):
		torch, dist = self.torch, self.dist
		(blob,) = blobs
		n = len(blob)
		h_in, h_out, d_in, d_out = self._staging(n)
		if n:
			h_in.numpy()[:n] = np.frombuffer(blob, dtype=np.uint8)
		if self.cuda:
			d_in.copy_(h_in, non_blocking=True)
			dist.all_gather_into_tensor(d_out, d_in)
			h_out.copy_(d_out, non_blocking=True)
			torch.cuda.current_stream().synchronize()
		else:
			dist.all_gather_into_tensor(d_out, d_in)
		raw = h_out.numpy()
		m = max(n, 1)
		return [raw[r * m:r * m + n].tobytes() for r in range(self.world)]

	def var(self, blobs):
		"""Blobs of different lengths in ONE collective: every rank sends a length header + its blob
		padded to a common capacity (grown, by all ranks alike, when some blob did not fit)."""
		(blob,) = blobs
		while True:
			cap = self._cap
			head = struct.pack("<q", len(blob))
			out = self([head + (blob[:cap] if len(blob) <= cap else b"") .ljust(cap, b"\0")])
			lens = [struct.unpack_from("<q", o, 0)[0] for o in out]
			if max(lens) <= cap:
				return [o[8:8 + k] for o, k in zip(out, lens)]
			self._cap = (max(lens) * 9 // 8 + 0xFFFF) & ~0xFFFF


def bind_to_gpu_numa_node(device_index):
	"""Pin this process to the CPUs of the NUMA node its GPU hangs off, so that pinned host buffers allocated
	afterwards (first touch) and the threads that fill them are local to the GPU's PCIe root: with 8 ranks streaming
	their audio shards at once, cross-socket copies are what limits the host-buffer path.  Returns a short
	description (or the reason nothing was done); never raises."""
	import os
	try:
		import pynvml as nv
		nv.nvmlInit()
		vis = os.environ.get("CUDA_VISIBLE_DEVICES")
		phys = device_index
		if vis:
			ids = [v.strip() for v in vis.split(",") if v.strip()]
			if device_index < len(ids) and ids[device_index].isdigit():
				phys = int(ids[device_index])
		bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(phys)).busId
		bus = bus.decode() if isinstance(bus, bytes) else bus
		bus = bus.lower()
		if len(bus.split(":")[0]) == 8:          # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
			bus = bus[4:]
		with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
			node = int(f.read().strip())
		if node < 0:
			return f"gpu {device_index} ({bus}): no NUMA affinity reported"
		with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
			spec = f.read().strip()
		cpus = set()
		for part in spec.split(","):
			a, _, b = part.partition("-")
			cpus.update(range(int(a), int(b or a) + 1))
		allowed = os.sched_getaffinity(0) & cpus
		if not allowed:
			return f"gpu {device_index} ({bus}): node {node} has no CPU this process may use"
		os.sched_setaffinity(0, allowed)
		return f"gpu {device_index} ({bus}): bound to NUMA node {node}, {len(allowed)} CPUs"
	except Exception as exc:                     # no NVML / no sysfs / not permitted: run unbound
		return f"gpu {device_index}: not bound ({type(exc).__name__}: {exc})"


class LinkedRun:
	"""One rank of a linked multi-GPU run: the hand-off, the bit tails and the packet records travel between the
	GPUs over NVLink peer memory inside one stream of kernels (csrc/link.cu); the host only launches and collects.
	`exchange(list of one blob) -> blobs of all ranks` is used once, to trade the IPC handles, and again only when a
	speculated slicer start state did not verify (then the repair protocol above takes over)."""

	def __init__(self, engine, rank, world, max_samples, exchange, exchange_var=None, tail_bits=16384):
		self.engine, self.rank, self.world = engine, rank, world
		self.exchange, self.exchange_var = exchange, exchange_var or exchange
		self.tail_bits = tail_bits if world > 1 else 0
		handle, _ = engine.link_create(rank, world, self.tail_bits, max_samples)
		# the link layout follows from (chains, world, tail_bits, max_samples): it has to be the same everywhere
		blobs = exchange([handle + struct.pack("<qqq", int(max_samples), int(self.tail_bits), engine.n_chains)])
		if len({b[64:] for b in blobs}) != 1:
			raise ValueError("LinkedRun: ranks disagree on max_samples / tail_bits / chain count")
		engine.link_connect(handles=[b[:64] for b in blobs])
		self.fallbacks = 0
		self.recoveries = 0

	def run(self, plan, audio_ptr, n_local, on_device=False, timing=None, fetch=True):
		"""-> (records, arena) of ALL ranks, ordered like an unsharded run.  fetch=False leaves them in the engine
		(host memory, like an unsharded pm_engine_run) and returns None when the fast path held."""
		import time
		t0 = time.perf_counter()
		self.engine.run_linked_begin(audio_ptr, n_local, plan, on_device)
		t1 = time.perf_counter()
		ok = self.engine.run_linked_end()
		t2 = time.perf_counter()
		if timing is not None:
			timing['linked_begin'] = timing.get('linked_begin', 0.0) + (t1 - t0) * 1e3
			timing['linked_end'] = timing.get('linked_end', 0.0) + (t2 - t1) * 1e3
		if ok == 1:
			return self.engine.fetch() if fetch else None
		if ok == 2:
			# some rank could not finish its decode from what it holds: all ranks decode the gathered bitstream
			self.recoveries += 1
			return recover_from_bitstream([self.engine], self.exchange_var)
		self.fallbacks += 1
		worker = ShardWorker(self.engine, plan, audio_ptr, n_local, on_device)
		return run_protocol([worker], self.exchange, self.exchange_var, timing=timing, resume=True)


def run_linked_local(demod_stack, audio, world, device=0, tail_bits=16384, **options):
	"""`world` ranks emulated on ONE GPU with the shard link (link buffers connected by plain device pointers):
	all ranks are enqueued first, then collected -- a waiting rank spins in a one-block kernel while the others
	run.  Returns (per-chain PacketMeta lists as rank 0 sees them, info)."""
	from .engine import Engine
	audio = np.ascontiguousarray(audio, dtype=np.int16)
	seg = int(options.get('segment_len', DEFAULT_SEGMENT_LEN))
	warm = int(options.get('warmup_len', DEFAULT_WARMUP_LEN))
	trim = max(_chain_trim(c) for c in demod_stack)
	sps = max(float(c[2].sample_rate) / float(c[2].symbol_rate) for c in demod_stack)
	plans = plan_shards(len(audio), world, segment_len=seg, warm_len=max(warm, seg), trim_max=trim,
		samples_per_symbol=sps, tail_bits=tail_bits)
	engines = [Engine(demod_stack, device=device, **options) for _ in plans]
	try:
		locals_ = [audio[p['audio_begin']:p['audio_end']] for p in plans]
		bases = [e.link_create(p['rank'], world, p['tail_bits'], max(len(l) for l in locals_))[1] for e, p in zip(engines, plans)]
		for e in engines:
			e.link_connect(pointers=bases)
		# One process, one GPU: cudaMalloc and first-use kernel loading synchronise the whole device, which would
		# block behind another rank's spinning wait kernel.  Size every buffer with a host-driven pass over the same
		# shards first.  (With one process per GPU nothing of the sort can happen: a rank only ever waits for *other*
		# devices.)
		run_protocol([ShardWorker(e, p, l.ctypes.data, len(l)) for e, p, l in zip(engines, plans, locals_)], local_exchange)
		for e, p, l in zip(engines, plans, locals_):
			e.run_linked_begin(l.ctypes.data, len(l), p)
		verified = [e.run_linked_end() for e in engines]
		info = dict(verified=verified, plans=plans, repairs=[e.stats()['slicer_repairs'] for e in engines])
		if all(v == 2 for v in verified):
			info['recovered'] = True
			return engines[0].packets(*recover_from_bitstream(engines, local_exchange)), info
		assert 2 not in verified, "ranks disagree on the recovery verdict"
		if all(verified):
			results = [e.fetch() for e in engines]
			info['all_ranks_equal'] = all(np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1])
				for r in results[1:])
			return engines[0].packets(*results[0]), info
		assert not any(verified), "ranks disagree on the hand-off verdict"
		workers = [ShardWorker(e, p, l.ctypes.data, len(l)) for e, p, l in zip(engines, plans, locals_)]
		timing = {}
		recs, arena = run_protocol(workers, local_exchange, resume=True, timing=timing)
		info['recovered'] = bool(timing.get('recovered'))
		return engines[0].packets(recs, arena), info
	finally:
		for e in engines:
			e.close()


def run_sharded_local(demod_stack, audio, world, device=0, tail_bits=16384, **options):
	"""Emulate `world` ranks on ONE GPU (one engine per shard, run in lock step).  Used by the
	GPU parity tests; the multi-GPU path (bench.py / run_sharded_distributed) runs the same
	ShardWorker protocol with one process per GPU."""
	from .engine import Engine
	audio = np.ascontiguousarray(audio, dtype=np.int16)
	seg = int(options.get('segment_len', DEFAULT_SEGMENT_LEN))
	warm = int(options.get('warmup_len', DEFAULT_WARMUP_LEN))
	trim = max(_chain_trim(c) for c in demod_stack)
	sps = max(float(c[2].sample_rate) / float(c[2].symbol_rate) for c in demod_stack)
	plans = plan_shards(len(audio), world, segment_len=seg, warm_len=max(warm, seg), trim_max=trim,
		samples_per_symbol=sps, tail_bits=tail_bits)
	engines = [Engine(demod_stack, device=device, **options) for _ in plans]
	try:
		workers = []
		for eng, plan in zip(engines, plans):
			local = audio[plan['audio_begin']:plan['audio_end']]
			workers.append(ShardWorker(eng, plan, local.ctypes.data, len(local)))
			workers[-1]._keep = local
		timing = {}
		recs, arena = run_protocol(workers, local_exchange, timing=timing)
		info = dict(rounds=[w.rounds for w in workers], plans=plans, recovered=bool(timing.get('recovered')),
			repairs=[e.stats()['slicer_repairs'] for e in engines])
		return engines[0].packets(recs, arena), info
	finally:
		for e in engines:
			e.close()


def _chain_trim(chain):
	modem = chain[1]
	if hasattr(modem, 'mark_correlator_i'):
		return (len(modem.input_bpf) - 1) + (len(modem.mark_correlator_i) - 1) + (len(modem.output_lpf) - 1)
	return len(modem.input_lpf) - 1
