"""Python face of the CUDA engine: chain objects -> pm_chain_desc table ->
pm_engine_run -> list[PacketMeta] per chain.

No CPU fallback: everything here calls libpymodem_b200.so through ctypes and
raises if the library or an sm_100 GPU is missing."""
import ctypes
import weakref
from collections.abc import Sequence

import numpy as np

from . import _lib
from .modems_codecs.packet_meta import PacketMeta


class EngineError(RuntimeError):
	pass


REC_DTYPE = np.dtype([
	('chain', '<u4'), ('len', '<u4'), ('offset', '<u8'), ('streamaddress', '<i8'),
	('bytes_corrected', '<u4'), ('calculated_crc', '<u2'), ('carried_crc', '<u2'),
	('valid_crc', 'u1'), ('valid_header', 'u1'), ('pad', 'u1', (6,)),
])
assert REC_DTYPE.itemsize == ctypes.sizeof(_lib.PacketRec) == 40


class PacketList(Sequence):
	"""The decode() result of one chain: a read-only sequence of PacketMeta in stream order, backed by the engine's
	record block (numpy structured array + byte arena).  Packets are built when they are first touched -- one hour of
	the 8-chain config is ~5400 records, and a Python object per record with a list-of-int payload costs more host time
	than the GPU needs for the whole decode -- so bulk consumers can stay on the arrays (.records, .arena) while
	code written against the reference's list[PacketMeta] (PacketMetaArray.add, iteration, indexing, len, ==) sees
	exactly that."""
	__slots__ = ('records', 'arena', 'name', '_raw', '_items')

	def __init__(self, records, arena, name, raw=None):
		self.records, self.arena, self.name = records, arena, name
		self._raw = raw
		self._items = None

	def __len__(self):
		return len(self.records)

	def _build(self):
		r = self.records
		raw = self._raw if self._raw is not None else self.arena.tobytes()
		name = self.name
		items = []
		for off, ln, sa, bc, calc, carried, vc, vh in zip(r['offset'].tolist(), r['len'].tolist(), r['streamaddress'].tolist(),
				r['bytes_corrected'].tolist(), r['calculated_crc'].tolist(), r['carried_crc'].tolist(),
				r['valid_crc'].tolist(), r['valid_header'].tolist()):
			p = PacketMeta.__new__(PacketMeta)
			p.__dict__ = {'data': list(raw[off:off + ln]), 'streamaddress': sa, 'source_sample_rate': 0.0,
				'CalculatedCRC': calc, 'CarriedCRC': carried, 'ValidCRC': bool(vc), 'SourceDecoder': name,
				'BytesCorrected': bc, 'CorrelatedDecoders': [], 'SlicedIQSamples': [], 'ValidHeader': bool(vh),
				'_device_checked': True}
			items.append(p)
		self._items = items
		return items

	def __getitem__(self, i):
		items = self._items if self._items is not None else self._build()
		return items[i]

	def __iter__(self):
		return iter(self._items if self._items is not None else self._build())

	def __eq__(self, other):
		if isinstance(other, (list, PacketList)):
			return len(self) == len(other) and all(a is b or a == b for a, b in zip(self, other))
		return NotImplemented

	def __repr__(self):
		return f"PacketList({self.name!r}, {len(self)} packets)"


def describe_chain(chain, keep):
	"""[object_name, modem, slicer, stream, codec] -> pm_chain_desc"""
	_name, modem, slicer, stream, codec = chain
	desc = _lib.ChainDesc()
	for block, what in ((modem, 'modem'), (slicer, 'slicer'), (stream, 'stream'), (codec, 'codec')):
		if block == [] or block is None:
			# the reference degrades a bad slicer/stream/codec spec to [] and the
			# chain then dies in its child process (pymodem.py:96-113)
			raise EngineError(f"chain '{_name}': invalid or missing '{what}'")
	modem.describe(desc, keep)
	slicer.describe(desc)
	stream.describe(desc)
	codec.describe(desc)
	return desc


class Engine:
	"""One engine = one GPU + one chain table (a demod_stack)."""

	def __init__(self, demod_stack, device=0, recordings=1, **options):
		"""recordings > 1: a batch engine -- the chain table is laid out `recordings` times, copy r decoding recording r of
		run_batch() (pm_chain_desc.recording); everything else about the engine is the same."""
		self._lib = _lib.load()
		self._h = ctypes.c_void_p()
		self.recordings = int(recordings)
		self.chains_per_recording = len(demod_stack)
		stack_in = demod_stack
		demod_stack = [chain for _ in range(self.recordings) for chain in stack_in]
		self.names = [chain[0] for chain in demod_stack]
		rc = self._lib.pm_engine_create(int(device), ctypes.byref(self._h))
		if rc != _lib.PM_OK:
			self._h = ctypes.c_void_p()
			raise EngineError(
				f"pm_engine_create(device={device}) failed ({rc}): a B200 (sm_100) GPU is required; "
				"there is no CPU fallback")
		for key, value in options.items():
			self.set_option(key, value)
		keep = []
		descs = (_lib.ChainDesc * len(demod_stack))()
		for i, chain in enumerate(demod_stack):
			descs[i] = describe_chain(chain, keep)
			descs[i].recording = i // self.chains_per_recording
		self.fingerprint = stack_fingerprint(stack_in)
		self._registered = {}
		self._check(self._lib.pm_engine_load_chains(self._h, descs, len(demod_stack)))
		self.n_chains = len(demod_stack)
		self.has_il2p = any(getattr(chain[4], 'codec_kind', None) == _lib.PM_CODEC_IL2P for chain in demod_stack)

	def _check(self, rc):
		if rc != _lib.PM_OK:
			msg = self._lib.pm_last_error(self._h)
			raise EngineError(f"pymodem_b200 error {rc}: {msg.decode() if msg else ''}")

	def close(self):
		if getattr(self, '_h', None) and self._h.value:
			self._lib.pm_engine_destroy(self._h)
			self._h = ctypes.c_void_p()

	def __del__(self):
		try:
			self.close()
		except Exception:
			pass

	def set_option(self, key, value):
		self._check(self._lib.pm_engine_set_option(self._h, key.encode(), float(value)))

	# -- execution -------------------------------------------------------------
	def run_raw(self, audio):
		"""int16 host ndarray -> (records structured array, arena uint8 array).
		This is the reference-facing C-ABI call with HOST buffers."""
		audio = np.ascontiguousarray(audio, dtype=np.int16)
		self._check(self._lib.pm_engine_run(self._h, audio.ctypes.data, audio.shape[0]))
		return self.fetch()

	def run_host_ptr(self, ptr, n):
		self._check(self._lib.pm_engine_run(self._h, ptr, n))

	def run_device_ptr(self, ptr, n):
		"""audio already resident in device memory (e.g. a torch int16 tensor's data_ptr())."""
		self._check(self._lib.pm_engine_run_device(self._h, ptr, n))

	# -- sharded execution (one recording split on the sample axis; see sharded.py) ------
	def shard_begin(self, audio_ptr, n, plan, on_device=False):
		"""plan: dict(sample_base, own_begin, own_len, first, last, tail_bits) -> ShardState array"""
		p = _lib.ShardPlan(int(plan['sample_base']), int(plan['own_begin']), int(plan['own_len']),
			int(bool(plan['first'])), int(bool(plan['last'])), int(plan['tail_bits']), int(plan.get('pre_segments', 0)))
		self._tail_words = int(plan['tail_bits']) // 32
		out = (_lib.ShardState * self.n_chains)()
		self._check(self._lib.pm_engine_shard_begin(self._h, audio_ptr, int(n), int(bool(on_device)), ctypes.byref(p), out))
		return out

	def shard_handoff(self, prev_states):
		"""prev_states: the previous shard's ShardState array (None on the first shard)
		-> (this shard's ShardState array, changed)"""
		out = (_lib.ShardState * self.n_chains)()
		changed = ctypes.c_int32(0)
		self._check(self._lib.pm_engine_shard_handoff(self._h, prev_states, out, ctypes.byref(changed)))
		return out, bool(changed.value)

	def shard_gather(self, symbols_before):
		"""symbols_before: per-chain symbols of all earlier shards -> this shard's tail (uint32[n_chains, tail_words])"""
		sb = (ctypes.c_int64 * self.n_chains)(*[int(x) for x in symbols_before])
		tail = np.zeros((self.n_chains, max(self._tail_words, 1)), dtype=np.uint32)
		self._check(self._lib.pm_engine_shard_gather(self._h, sb, tail.ctypes.data))
		return tail[:, :self._tail_words]

	def shard_finish(self, tail_in):
		if tail_in is None:
			self._check(self._lib.pm_engine_shard_finish(self._h, None))
		else:
			t = np.ascontiguousarray(tail_in, dtype=np.uint32)
			self._check(self._lib.pm_engine_shard_finish(self._h, t.ctypes.data))

	def shard_finish_il2p(self, tail_in, prev_states):
		"""Finish of a shard whose engine has IL2P chains: needs the previous shard's decoder states (None on the
		first shard) and returns this shard's (an Il2pState array) for the next one."""
		out = (_lib.Il2pState * self.n_chains)()
		t = None if tail_in is None else np.ascontiguousarray(tail_in, dtype=np.uint32)
		self._check(self._lib.pm_engine_shard_finish_il2p(self._h, None if t is None else t.ctypes.data, prev_states, out))
		return out

	# -- shard link: the hand-off done by the GPUs over peer memory (csrc/link.cu) ---------------
	def link_create(self, rank, world, tail_bits, max_samples):
		"""-> (64-byte CUDA IPC handle of this rank's link buffer, its device address)"""
		handle = ctypes.create_string_buffer(64)
		base = ctypes.c_void_p()
		self._check(self._lib.pm_engine_link_create(self._h, int(rank), int(world), int(tail_bits), int(max_samples),
			handle, ctypes.byref(base)))
		return handle.raw, int(base.value)

	def link_connect(self, handles=None, pointers=None):
		"""handles: the IPC handles of all ranks in rank order (one process per GPU); pointers: the link buffer
		addresses of all ranks (several engines inside one process)."""
		if handles is not None:
			blob = b"".join(handles)
			self._check(self._lib.pm_engine_link_connect(self._h, ctypes.c_char_p(blob), 1))
		else:
			arr = (ctypes.c_void_p * len(pointers))(*pointers)
			self._check(self._lib.pm_engine_link_connect(self._h, arr, 0))

	def run_linked_begin(self, audio_ptr, n, plan, on_device=False):
		p = _lib.ShardPlan(int(plan['sample_base']), int(plan['own_begin']), int(plan['own_len']),
			int(bool(plan['first'])), int(bool(plan['last'])), int(plan['tail_bits']), int(plan.get('pre_segments', 0)))
		self._tail_words = int(plan['tail_bits']) // 32
		self._check(self._lib.pm_engine_run_linked_begin(self._h, audio_ptr, int(n), int(bool(on_device)), ctypes.byref(p)))

	def run_linked_end(self):
		"""-> 1 when every hand-off verified (merged records of all ranks are ready for fetch()); 0 when the ranks have
		to fall back to the host-driven slicer repair protocol; 2 when some rank could not finish its decode from what
		it holds and all ranks recover from the gathered bitstream (sharded.recover_from_bitstream)."""
		v = ctypes.c_int32(0)
		self._check(self._lib.pm_engine_run_linked_end(self._h, ctypes.byref(v)))
		return int(v.value)

	def shard_states(self):
		out = (_lib.ShardState * self.n_chains)()
		self._check(self._lib.pm_engine_shard_states(self._h, out))
		return out

	def fetch(self):
		n = self._lib.pm_engine_num_packets(self._h)
		nb = self._lib.pm_engine_arena_bytes(self._h)
		if n < 0:
			raise EngineError("no results: run the engine first")
		recs = np.empty(n, dtype=REC_DTYPE)
		arena = np.empty(max(nb, 1), dtype=np.uint8)
		self._check(self._lib.pm_engine_get_packets(self._h, recs.ctypes.data, n, arena.ctypes.data, arena.shape[0]))
		return recs, arena[:nb]

	def packets(self, recs, arena):
		"""records -> one PacketList (a sequence of PacketMeta) per chain, in config order (the order the reference's
		deterministic driver produces).  The records are ordered by (chain, stream position), so every chain is a
		contiguous block: no per-record work happens here."""
		bounds = np.searchsorted(recs['chain'], np.arange(self.n_chains + 1, dtype=np.uint32))
		raw = arena.tobytes()
		return [PacketList(recs[bounds[c]:bounds[c + 1]], arena, self.names[c], raw) for c in range(self.n_chains)]

	# -- host memory ---------------------------------------------------------------
	def pin(self, audio):
		"""Page-lock a NumPy buffer in place (cudaHostRegister), once: every later run() on it is DMA'ed straight from
		the caller's memory.  The registration is dropped when the array is garbage collected."""
		key = (audio.ctypes.data, audio.nbytes)
		if key in self._registered or audio.nbytes == 0:
			return True
		if self._lib.pm_host_register(audio.ctypes.data, audio.nbytes) != _lib.PM_OK:
			return False
		self._registered[key] = weakref.finalize(audio, _unregister, self._lib, audio.ctypes.data, self._registered, key)
		return True

	# -- the reference's blocks one at a time (chain_execute.py:32-47) -------------
	def slice_soft(self, chain, soft_i, soft_q=None):
		"""slicer.slice(): float64 soft values -> (bytes uint8[n], addresses int64[n]) of the AddressedData stream."""
		i = np.ascontiguousarray(soft_i, dtype=np.float64)
		q = None if soft_q is None else np.ascontiguousarray(soft_q, dtype=np.float64)
		if q is not None and len(q) != len(i):
			raise EngineError("slice_soft: I and Q lengths differ")
		self._check(self._lib.pm_engine_slice_soft(self._h, chain, i.ctypes.data, None if q is None else q.ctypes.data, len(i)))
		return self.stream(chain, 0)

	def unscramble_stream(self, chain, data, addresses):
		"""stream.stream_unscramble_8bit(): (bytes, addresses) -> (bytes, addresses)"""
		b = np.ascontiguousarray(data, dtype=np.uint8)
		a = np.ascontiguousarray(addresses, dtype=np.int64)
		self._check(self._lib.pm_engine_unscramble_stream(self._h, chain, b.ctypes.data, a.ctypes.data, len(b)))
		return self.stream(chain, 1)

	def decode_stream(self, chain, data, addresses):
		"""codec.decode(): (bytes, addresses) -> (records, arena)"""
		b = np.ascontiguousarray(data, dtype=np.uint8)
		a = np.ascontiguousarray(addresses, dtype=np.int64)
		self._check(self._lib.pm_engine_decode_stream(self._h, chain, b.ctypes.data, a.ctypes.data, len(b)))
		return self.fetch()

	def shard_export(self, chain):
		"""What this shard holds of a chain's sliced stream after shard_gather: (bit words uint32[], byte addresses
		uint32[], info = [local bits, first own bit, own bits, sample_base])."""
		info = (ctypes.c_int64 * 4)()
		self._check(self._lib.pm_engine_shard_export(self._h, chain, None, 0, None, 0, info))
		nw, nby = (info[0] + 31) // 32, (info[0] + 7) // 8
		bits = np.zeros(max(nw, 1), dtype=np.uint32)
		addr = np.zeros(max(nby, 1), dtype=np.uint32)
		self._check(self._lib.pm_engine_shard_export(self._h, chain, bits.ctypes.data, len(bits), addr.ctypes.data, len(addr), info))
		return bits[:nw], addr[:nby], [int(x) for x in info]

	def signs(self, chain, component=0):
		"""Packed signs of the soft values as the slicer reads them: uint32[(soft_len + 31) // 32]."""
		n = self._lib.pm_engine_soft_len(self._h, chain)
		if n < 0:
			raise EngineError("no signs: run the engine first")
		out = np.zeros(max((n + 31) // 32, 1), dtype=np.uint32)
		self._check(self._lib.pm_engine_get_signs(self._h, chain, component, out.ctypes.data, len(out)))
		return out[:(n + 31) // 32]

	def run(self, audio):
		recs, arena = self.run_raw(audio)
		return self.packets(recs, arena)

	def run_batch(self, audios):
		"""One engine call for len(audios) == recordings recordings (int16 arrays, lengths may differ):
		-> [per recording [per chain PacketList]].  All stages run over every chain of every recording at once."""
		if len(audios) != self.recordings:
			raise EngineError(f"run_batch: this engine was built for {self.recordings} recordings, got {len(audios)}")
		if self.recordings == 1:
			# a batch of one is a plain run: straight from the caller's memory, the copy overlapped with the front end,
			# instead of a pass through the batch's staging buffer (3600 s of afsk_1200.json: 51 -> ~20 ms per call)
			return [self.run(audios[0])]
		lens = np.array([len(a) for a in audios], dtype=np.int64)
		stride = int(max(int(lens.max()), 1) + 63) // 64 * 64
		buf = getattr(self, '_batch_buf', None)
		if buf is None or buf.shape != (self.recordings, stride):
			buf = self._batch_buf = pinned_empty(self.recordings * stride).reshape(self.recordings, stride)
		for r, a in enumerate(audios):
			buf[r, :len(a)] = a
		self._check(self._lib.pm_engine_run_batch(self._h, buf.ctypes.data, stride,
			lens.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), self.recordings))
		per_chain = self.packets(*self.fetch())
		c = self.chains_per_recording
		return [per_chain[r * c:(r + 1) * c] for r in range(self.recordings)]

	# -- intermediates (parity tests) -------------------------------------------
	def soft(self, chain, component=0):
		n = self._lib.pm_engine_soft_len(self._h, chain)
		if n < 0:
			raise EngineError("no soft values: run with keep_soft=1 first")
		out = np.zeros(n, dtype=np.float32)
		self._check(self._lib.pm_engine_get_soft(self._h, chain, component, out.ctypes.data, n))
		return out

	def stream(self, chain, stage):
		"""AddressedData stream of a chain: stage 0 = slicer output, 1 = after LFSR.
		-> (bytes uint8[n], addresses int64[n])"""
		n = self._lib.pm_engine_stream_len(self._h, chain)
		if n < 0:
			raise EngineError("no stream: run the engine first")
		b = np.zeros(max(n, 1), dtype=np.uint8)
		a = np.zeros(max(n, 1), dtype=np.int64)
		self._check(self._lib.pm_engine_get_stream(self._h, chain, stage, b.ctypes.data, a.ctypes.data, n))
		return b[:n], a[:n]

	def stats(self):
		s = _lib.Stats()
		self._check(self._lib.pm_engine_get_stats(self._h, ctypes.byref(s)))
		return s.as_dict()

	def kernel_times(self):
		"""Per-kernel times of the last run (option kernel_times=1): [(name, launches, ms)] in order of first launch."""
		buf = ctypes.create_string_buffer(1 << 16)
		n = self._lib.pm_engine_kernel_times(self._h, buf, len(buf))
		if n < 0:
			raise EngineError(f"pm_engine_kernel_times failed ({n})")
		out = []
		for line in buf.value.decode().splitlines():
			name, cnt, ms = line.split("\t")
			out.append((name, int(cnt), float(ms)))
		return out

	def trace(self):
		"""Timeline of the last host-buffer run (option trace=1): [(label, ms since the start of the run)]."""
		buf = ctypes.create_string_buffer(1 << 16)
		if self._lib.pm_engine_trace(self._h, buf, len(buf)) < 0:
			raise EngineError("pm_engine_trace failed")
		return [(l.rsplit(" ", 1)[0], float(l.rsplit(" ", 1)[1])) for l in buf.value.decode().splitlines()]

	def stage_clocks(self):
		"""Cycles per front-end stage summed over CTAs since the last call (option stage_clocks=1); see the header."""
		out = (ctypes.c_uint64 * 8)()
		self._check(self._lib.pm_engine_stage_clocks(self._h, out))
		return list(out)

	def front_macs_per_sample(self):
		return self._lib.pm_engine_front_macs_per_sample(self._h)

	def front_tensor_macs_per_sample(self):
		return self._lib.pm_engine_front_tensor_macs_per_sample(self._h)


def _unregister(lib, ptr, table, key):
	table.pop(key, None)
	lib.pm_host_unregister(ptr)


_SCALARS = (int, float, str, bool, type(None))


def _state_key(obj, depth=0):
	"""Everything a block object holds, as a hashable tuple: arrays by content (dtype, shape, hash of the bytes), scalars
	as they are, plain lists as tuples, nested helper objects (AGC, NCO, IIR_1, PI_control, RRC, Hilbert ...) recursively.
	(Python's own hashing instead of a digest over repr()s: 0.35 -> 0.1 ms per process_chains call for the eight
	chains of the headline config -- that call is inside the end-to-end figure.)"""
	d = vars(obj)
	vals = []
	for v in d.values():
		t = type(v)
		if t in _SCALARS:
			vals.append(v)
		elif t is np.ndarray:
			vals.append((v.dtype.char, v.shape, hash(v.tobytes())))
		elif t in (list, tuple):
			vals.append(tuple(v) if (not v or type(v[0]) in _SCALARS) else repr(v))
		elif hasattr(v, '__dict__') and depth < 4:
			vals.append(_state_key(v, depth + 1))
		else:
			vals.append(repr(v))
	return (type(obj).__name__, tuple(d), tuple(vals))


def stack_fingerprint(demod_stack):
	"""The complete state of every block of every chain (parameters, tap and table arrays) as one hashable key.  The
	reference reads its blocks' state on every call (chain_execute.py:32-47), so a retune() or StringOptionsRetune()
	between two calls must take effect: engine_for() compares this, not object identities.  (Any attribute change gives
	a new engine, also one describe() would not look at: conservative, never stale.  float('nan') attributes compare
	unequal to themselves and so only cost a rebuild.)"""
	out = []
	for chain in demod_stack:
		blocks = [str(chain[0])]
		for block in chain[1:]:
			blocks.append("<missing>" if (block == [] or block is None) else _state_key(block))
		out.append(tuple(blocks))
	return tuple(out)


_cache = {}


def engine_for(demod_stack, device=0, recordings=1, **options):
	"""One cached engine per (chain parameters, device, recordings, options): building an engine uploads taps and allocates
	its buffers.  The key is the digest of the described parameters, so retuned blocks get a new engine."""
	key = (stack_fingerprint(demod_stack), device, recordings, tuple(sorted(options.items())))
	eng = _cache.get(key)
	if eng is None:
		eng = Engine(demod_stack, device=device, recordings=recordings, **options)
		for old in _cache.values():
			old.close()
		_cache.clear()
		_cache[key] = eng
	return eng


class _NoModem:
	"""Stands in for the modem of a chain that only serves the per-stage calls."""
	modem_kind = _lib.PM_MODEM_NONE

	def describe(self, desc, keep):
		desc.modem_kind = self.modem_kind


def stage_engine(slicer=None, stream=None, codec=None, name="stage"):
	"""A one-chain engine around the block(s) a per-stage call needs (the others are defaults that stay unused)."""
	from .modems_codecs import ax25, lfsr, slicer as slicer_mod
	chain = [name, _NoModem(), slicer or slicer_mod.BinarySlicer(sample_rate=48000, config='1200'), stream or lfsr.LFSR(),
		codec or ax25.AX25Codec(ident=name)]
	return engine_for([chain])


def addressed_arrays(stream):
	"""list[AddressedData] (or a (bytes, addresses) pair of arrays) -> (uint8[n], int64[n])"""
	if isinstance(stream, tuple) and len(stream) == 2:
		return np.ascontiguousarray(stream[0], dtype=np.uint8), np.ascontiguousarray(stream[1], dtype=np.int64)
	n = len(stream)
	return (np.fromiter((int(s.data) & 0xFF for s in stream), dtype=np.uint8, count=n),
		np.fromiter((int(s.address) for s in stream), dtype=np.int64, count=n))


def addressed_list(data, addresses):
	from .modems_codecs.data_classes import AddressedData
	return [AddressedData(d, a) for d, a in zip(data.tolist(), addresses.tolist())]


def pinned_empty(n, dtype=np.int16):
	"""A NumPy array in page-locked host memory (pm_host_alloc): pm_engine_run DMAs straight out of it.  Read a
	recording into one of these (python -m pymodem_b200 does, pymodem.py:46) instead of into pageable memory.
	The memory is released when the last view of it goes away."""
	lib = _lib.load()
	dtype = np.dtype(dtype)
	nbytes = max(int(n) * dtype.itemsize, 1)
	ptr = lib.pm_host_alloc(nbytes)
	if not ptr:
		raise EngineError(f"pm_host_alloc({nbytes}) failed: a CUDA device is required")
	buf = (ctypes.c_char * nbytes).from_address(ptr)
	weakref.finalize(buf, lib.pm_host_free, ptr)
	return np.frombuffer(buf, dtype=dtype, count=int(n))


def read_wav_pinned(path):
	"""scipy.io.wavfile.read (pymodem.py:46) with the samples landing in pinned memory: the file is memory-mapped and
	copied once, straight into the buffer the GPU reads.  -> (sample_rate, int16 ndarray)"""
	from scipy.io.wavfile import read as readwav
	rate, mapped = readwav(path, mmap=True)
	if mapped.ndim != 1 or mapped.dtype != np.int16:
		raise ValueError("16-bit mono PCM expected")
	out = pinned_empty(len(mapped), np.int16)
	out[:] = mapped
	return int(rate), out


def demod_only(modem, audio):
	"""modem.demod(audio) on the GPU: float64 ndarray of soft values (IQData for MPSKModem, psk.py:748)."""
	from . import _lib as L
	from .modems_codecs import ax25, data_classes, lfsr, slicer
	rate = getattr(modem, 'output_sample_rate', modem.sample_rate)
	iq = modem.modem_kind == L.PM_MODEM_MPSK
	if iq:
		s = slicer.QuadratureSlicer(sample_rate=rate, config='qpsk_2400')
		s.retune(symbol_rate=modem.symbol_rate)
	else:
		s = slicer.BinarySlicer(sample_rate=rate, config='1200')
		s.retune(symbol_rate=getattr(modem, 'symbol_rate', 1200))
	eng = Engine([["demod", modem, s, lfsr.LFSR(), ax25.AX25Codec()]], keep_soft=1)
	try:
		eng.run_raw(audio)
		if not iq:
			return eng.soft(0).astype(np.float64)
		out = data_classes.IQData()
		out.i_data = eng.soft(0, 0).astype(np.float64)
		out.q_data = eng.soft(0, 1).astype(np.float64)
		return out
	finally:
		eng.close()


def measure_fp64_chain(device=0):
	"""(nanoseconds, SM cycles) per dependent float64 operation: the floor under the sequential carrier loops."""
	ns, cyc = ctypes.c_double(), ctypes.c_double()
	rc = _lib.load().pm_measure_fp64_chain(device, ctypes.byref(ns), ctypes.byref(cyc))
	if rc != _lib.PM_OK:
		raise EngineError(f"pm_measure_fp64_chain failed ({rc})")
	return ns.value, cyc.value


def measure_fp32_peak(device=0):
	t = ctypes.c_double()
	rc = _lib.load().pm_measure_fp32_peak(device, ctypes.byref(t))
	if rc != _lib.PM_OK:
		raise EngineError(f"pm_measure_fp32_peak failed ({rc})")
	return t.value
