"""CPU stand-in for pymodem_b200.engine.Engine's shard_* methods, built on the oracle.

TEST INFRASTRUCTURE: it exists so that the host-side shard protocol
(pymodem_b200/sharded.py: planning, hand-off rounds, symbol-count scan, bit tails, record
merge) can be exercised without a GPU -- in one process and over torch.distributed/gloo
with world_size 2.  It follows the semantics documented in include/pymodem_b200.h
(pm_engine_shard_*); the CUDA engine is tested against the same expectations on the GPU
(tests/test_gpu_sharded.py)."""
import ctypes
import struct

import numpy as np

from oracle import oracle as orc
from pymodem_b200 import _lib
from pymodem_b200.engine import REC_DTYPE


def _bits_of(x):
	return struct.unpack("<q", struct.pack("<d", x))[0]


class _Chain:
	def __init__(self, sample_rate, line):
		self.c = orc.Chain(sample_rate, line)
		self.lock = self.c.slicer.lock_rate
		self.symbol_rate = self.c.slicer.symbol_rate
		self.rate = self.c.slicer.sample_rate

	def slicer_state(self, clock, last_nonneg, streamaddress=0, bit_count=0):
		st = orc._Slicer()
		orc.lib().orc_slicer_init(ctypes.byref(st), float(self.rate), float(self.symbol_rate), float(self.lock))
		st.phase_clock = clock
		st.last_sample = 0.0 if last_nonneg else -1.0
		st.streamaddress = streamaddress
		st.working_bit_count = bit_count
		st.working_byte = 0
		return st

	@staticmethod
	def run(st, soft):
		soft = np.ascontiguousarray(soft, dtype=np.float64)
		cap = len(soft) // 2 + 16
		b = np.empty(cap, dtype=np.uint8)
		a = np.empty(cap, dtype=np.int64)
		n = orc.lib().orc_binary_slice(ctypes.byref(st), soft.ctypes.data_as(ctypes.c_void_p), len(soft),
			b.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p), cap)
		return b[:n], a[:n]


def _ax25_bits(bits, valid_from=0):
	"""ax25.py:25-93 over bits[valid_from:], reporting per flag (position, data bytes or None, whether a
	reliable flag -- all 8 pattern bits valid -- came before, whether an abort was seen since the start)."""
	out = []
	wb = one = bit_index = byte_index = 0
	data = []
	first_flag_seen, abort_before_first = False, False
	for g, bit in enumerate(bits):
		if g < valid_from:
			continue
		if bit:
			wb |= 0x80
			one += 1
			bit_index += 1
			if one > 6:
				bit_index = 0
				byte_index = 0
				if not first_flag_seen:
					abort_before_first = True
			if bit_index == 8:
				bit_index = 0
				data.append(wb)
				byte_index += 1
				if byte_index > 1023:
					byte_index = 0
					one = 0
			wb >>= 1
		else:
			if one < 5:
				bit_index += 1
				if bit_index == 8:
					bit_index = 0
					data.append(wb)
					byte_index += 1
					if byte_index > 1023:
						byte_index = 0
				wb >>= 1
			elif one == 6:
				emit = byte_index >= 18 and bit_index == 7
				out.append((g, bytes(data) if emit else None, first_flag_seen, abort_before_first))
				if g >= valid_from + 8:
					first_flag_seen = True
				data = []
				byte_index = 0
				bit_index = 0
			one = 0
	return out


class SimEngine:
	def __init__(self, sample_rate, lines, warm_len=8192):
		self.chains = [_Chain(sample_rate, l) for l in lines if l.get('object_type') == 'demod_chain']
		self.n_chains = len(self.chains)
		self.warm_len = warm_len

	# -- pm_engine_shard_begin -------------------------------------------------------------
	def shard_begin(self, audio, n, plan, on_device=False):
		self.plan = plan
		audio = np.asarray(audio[:n])
		self.soft = [ch.c.modem.demod(audio) for ch in self.chains]
		self.start = []
		for ch, soft in zip(self.chains, self.soft):
			if plan['first']:
				self.start.append((0.0, True))
			else:
				st = ch.slicer_state(0.0, True)
				w0 = max(0, plan['own_begin'] - self.warm_len)
				ch.run(st, soft[w0:plan['own_begin']])
				self.start.append((st.phase_clock, st.last_sample >= 0))
		return self._run_own()

	def _own_end(self, soft):
		return len(soft) if self.plan['last'] else min(len(soft), self.plan['own_begin'] + self.plan['own_len'])

	def _run_own(self):
		out = (_lib.ShardState * self.n_chains)()
		for i, (ch, soft) in enumerate(zip(self.chains, self.soft)):
			st = ch.slicer_state(self.start[i][0], self.start[i][1])
			b, _ = ch.run(st, soft[self.plan['own_begin']:self._own_end(soft)])
			out[i].start_clock, out[i].start_last, out[i].start_last_q = self.start[i][0], int(self.start[i][1]), 1
			out[i].end_clock, out[i].end_last, out[i].end_last_q = st.phase_clock, int(st.last_sample >= 0), 1
			out[i].n_symbols = 8 * len(b) + st.working_bit_count
		self.states = out
		return out

	# -- pm_engine_shard_handoff -----------------------------------------------------------
	def shard_handoff(self, prev):
		changed = False
		if not self.plan['first']:
			want = [(prev[i].end_clock, bool(prev[i].end_last)) for i in range(self.n_chains)]
			if any(_bits_of(w[0]) != _bits_of(s[0]) or w[1] != s[1] for w, s in zip(want, self.start)):
				before = [(_bits_of(s.end_clock), s.end_last, s.n_symbols) for s in self.states]
				self.start = want
				self._run_own()
				changed = before != [(_bits_of(s.end_clock), s.end_last, s.n_symbols) for s in self.states]
		return self.states, changed

	# -- pm_engine_shard_gather ------------------------------------------------------------
	def shard_gather(self, symbols_before):
		K = self.plan['tail_bits']
		self.own_bits, self.own_addr, self.P = [], [], list(symbols_before)
		tails = np.zeros((self.n_chains, max(K // 32, 1)), dtype=np.uint32)
		for i, (ch, soft) in enumerate(zip(self.chains, self.soft)):
			P = 0 if self.plan['first'] else int(symbols_before[i])
			a = P % 8
			st = ch.slicer_state(self.start[i][0], self.start[i][1], streamaddress=self.plan['own_begin'], bit_count=a)
			b, addr = ch.run(st, soft[self.plan['own_begin']:])
			bits = np.unpackbits(b)[a:]                    # own (+ forward) bits, whole bytes only
			self.own_bits.append(bits)
			self.own_addr.append(addr)                    # address of global byte P//8 + j
			if not self.plan['last'] and K:
				n_own = int(self.states[i].n_symbols)
				if n_own < K:
					raise RuntimeError("shard holds fewer bits than the hand-off tail")
				t = bits[n_own - K:n_own]
				tails[i, :K // 32] = np.packbits(t.reshape(-1, 32), axis=1, bitorder='little').view(np.uint32).ravel()
		return tails[:, :K // 32]

	# -- pm_engine_shard_finish ------------------------------------------------------------
	def shard_finish(self, tail_in):
		K = self.plan['tail_bits']
		recs, arena = [], bytearray()
		for i, ch in enumerate(self.chains):
			valid_from = 0
			if self.plan['first']:
				P, A0, bit_off, tail = 0, 0, 0, np.zeros(0, dtype=np.uint8)
			else:
				P = int(self.P[i])
				A0 = ((P - K) >> 3) << 3
				bit_off = P - A0
				tail = np.unpackbits(np.asarray(tail_in[i], dtype=np.uint32).view(np.uint8), bitorder='little')
				valid_from = bit_off - K + max(ch.c.stream.polynomial.bit_length() - 1, 0)
			stream = np.concatenate([np.zeros(bit_off - len(tail), dtype=np.uint8), tail, self.own_bits[i]])
			stream = stream[:len(stream) // 8 * 8]
			by = np.packbits(stream)
			dby, _ = ch.c.stream.__class__({'poly': hex(ch.c.stream.polynomial), 'invert': str(ch.c.stream.invert)}) \
				.stream_unscramble_8bit(by, None)
			dbits = np.unpackbits(dby)
			n_own = int(self.states[i].n_symbols)
			own_lo = bit_off
			own_hi = len(dbits) + 1 if self.plan['last'] else bit_off + n_own
			first_byte_global = P // 8 - A0 // 8           # local byte index of the first own byte
			for (g, data, flag_before, abort_before) in _ax25_bits(dbits, valid_from):
				if not (own_lo <= g < own_hi):
					continue
				if not flag_before and not self.plan['first'] and (data is not None or not abort_before):
					raise RuntimeError("a frame reaches back past the hand-off tail")
				if data is None:
					continue
				j = (g >> 3) - first_byte_global
				addr = int(self.own_addr[i][j]) + self.plan['sample_base']
				carried, calc, valid = orc.check_crc(data)
				recs.append((i, len(data), len(arena), addr, 0, calc, carried, int(valid), int(orc.validate_header(data)),
					[0] * 6))
				arena += data
		self._recs = np.array(recs, dtype=REC_DTYPE) if recs else np.zeros(0, dtype=REC_DTYPE)
		self._arena = np.frombuffer(bytes(arena), dtype=np.uint8)

	def fetch(self):
		return self._recs, self._arena
