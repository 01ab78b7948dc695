"""Host side of the multi-GPU path on CPU: shard planning, the hand-off protocol, symbol-count
scan, bit tails and record merge (pymodem_b200/sharded.py) -- driven with the oracle-backed
SimEngine in one process and over torch.distributed (gloo, world_size 2)."""
import os
import sys

import numpy as np
import pytest

from util import Golden

HERE = os.path.dirname(os.path.abspath(__file__))


def _tuples(recs, arena, n_chains):
	out = [[] for _ in range(n_chains)]
	raw = arena.tobytes()
	for r in recs:
		out[int(r['chain'])].append((int(r['streamaddress']), raw[int(r['offset']):int(r['offset']) + int(r['len'])],
			int(r['bytes_corrected'])))
	return out


def test_plan_shards_covers_the_recording():
	from pymodem_b200.sharded import plan_shards
	n = 172_800_000
	for world in (1, 2, 4, 8):
		plans = plan_shards(n, world, segment_len=16384, warm_len=16384, trim_max=305)
		assert len(plans) == world and plans[0]['first'] and plans[-1]['last']
		assert plans[0]['sample_base'] == 0 and plans[0]['own_begin'] == 0
		for a, b in zip(plans, plans[1:]):
			# own ranges tile the sample axis without gaps
			assert a['sample_base'] + a['own_begin'] + a['own_len'] == b['sample_base'] + b['own_begin']
			assert b['own_begin'] % 16384 == 0 and a['own_len'] % 16384 == 0
			assert b['audio_begin'] <= b['sample_base'] + b['own_begin'] - 16384      # warm-up history
			assert a['audio_end'] >= a['sample_base'] + a['own_begin'] + a['own_len'] + 305   # FIR halo + forward symbols
		assert plans[-1]['audio_end'] == n
	with pytest.raises(ValueError):
		plan_shards(40000, 4, segment_len=16384)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_protocol_in_one_process_matches_unsharded(oracle, world):
	from pymodem_b200.sharded import ShardWorker, local_exchange, plan_shards, run_protocol
	from shard_sim import SimEngine
	g = Golden("afsk1200_superopt_48k")
	audio = g.audio()
	lines = g.chain_lines()[:3]
	plans = plan_shards(len(audio), world, segment_len=4096, warm_len=8192, trim_max=305, tail_bits=2048)
	workers = []
	for plan in plans:
		local = audio[plan['audio_begin']:plan['audio_end']]
		workers.append(ShardWorker(SimEngine(g.sample_rate, lines, warm_len=8192), plan, local, len(local)))
	recs, arena = run_protocol(workers, local_exchange)
	assert _tuples(recs, arena, len(lines)) == g.all_packets()[:3]


def test_wrong_speculation_is_repaired_across_rounds(oracle):
	"""No warm-up at all: every rank's speculated start state is wrong, the hand-off must repair it
	(and the repair of rank r changes what rank r+1 has to verify against)."""
	from pymodem_b200.sharded import ShardWorker, local_exchange, plan_shards, run_protocol
	from shard_sim import SimEngine
	g = Golden("afsk1200_superopt_48k")
	audio = g.audio()
	lines = g.chain_lines()[:2]
	plans = plan_shards(len(audio), 4, segment_len=4096, warm_len=4096, trim_max=305, tail_bits=2048)
	workers = []
	for plan in plans:
		local = audio[plan['audio_begin']:plan['audio_end']]
		workers.append(ShardWorker(SimEngine(g.sample_rate, lines, warm_len=0), plan, local, len(local)))
	recs, arena = run_protocol(workers, local_exchange)
	assert _tuples(recs, arena, len(lines)) == g.all_packets()[:2]


def test_silence_needs_one_round_per_rank(oracle):
	"""Digital silence has no zero crossings, so a wrong start state never merges with the true
	trajectory: the repair of rank r changes its end state and rank r+1 has to be repaired in the
	next round.  The symbol counts must still add up to the sequential slicer's."""
	from pymodem_b200 import configs
	from pymodem_b200.sharded import ShardWorker, local_exchange, plan_shards, run_protocol
	from shard_sim import SimEngine
	lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())[:1]
	audio = np.zeros(48000 * 6 + 1234, dtype=np.int16)
	world = 4
	plans = plan_shards(len(audio), world, segment_len=4096, warm_len=4096, trim_max=305, tail_bits=64)
	workers = []
	for plan in plans:
		local = audio[plan['audio_begin']:plan['audio_end']]
		workers.append(ShardWorker(SimEngine(48000, lines, warm_len=1000), plan, local, len(local)))
	recs, _ = run_protocol(workers, local_exchange)
	assert len(recs) == 0
	assert max(w.rounds for w in workers) >= world - 1
	chain = oracle.Chain(48000, lines[0])
	b, _ = chain.slicer.slice(chain.modem.demod(audio))
	total = sum(int(w.states[0].n_symbols) for w in workers)
	assert total // 8 == len(b)


def _gloo_worker(rank, world, port, tag, ret):
	sys.path.insert(0, HERE)
	sys.path.insert(0, os.path.dirname(HERE))
	import torch.distributed as dist
	from pymodem_b200.sharded import ShardWorker, TorchExchange, plan_shards, run_protocol
	from shard_sim import SimEngine
	os.environ["MASTER_ADDR"] = "127.0.0.1"
	os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	try:
		g = Golden(tag)
		audio = g.audio()
		lines = g.chain_lines()[:2]
		plan = plan_shards(len(audio), world, segment_len=4096, warm_len=8192, trim_max=305, tail_bits=2048)[rank]
		local = audio[plan['audio_begin']:plan['audio_end']]
		worker = ShardWorker(SimEngine(g.sample_rate, lines), plan, local, len(local))
		ex = TorchExchange("cpu")
		recs, arena = run_protocol([worker], ex, ex.var)
		ret[rank] = _tuples(recs, arena, len(lines))
	finally:
		dist.destroy_process_group()


def test_protocol_over_gloo_world2(oracle):
	import torch.multiprocessing as mp
	g = Golden("afsk1200_superopt_48k")
	ctx = mp.get_context("spawn")
	with ctx.Manager() as mgr:
		ret = mgr.dict()
		port = 29500 + os.getpid() % 2000
		procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, "afsk1200_superopt_48k", ret)) for r in range(2)]
		for p in procs:
			p.start()
		for p in procs:
			p.join(timeout=240)
		assert all(p.exitcode == 0 for p in procs)
		want = g.all_packets()[:2]
		assert ret[0] == want and ret[1] == want      # every rank ends up with the merged, ordered record set


def test_merge_results_orders_by_chain_then_rank():
	import struct
	from pymodem_b200.engine import REC_DTYPE
	from pymodem_b200.sharded import merge_results

	def blob(rows, arena):
		recs = np.array(rows, dtype=REC_DTYPE)
		return struct.pack("<qq", len(recs), len(arena)) + recs.tobytes() + bytes(arena)
	z = [0] * 6
	b0 = blob([(0, 2, 0, 10, 0, 0, 0, 0, 0, z), (1, 1, 2, 20, 0, 0, 0, 0, 0, z)], b"\x01\x02\x03")
	b1 = blob([(0, 1, 0, 30, 0, 0, 0, 0, 0, z)], b"\x04")
	recs, arena = merge_results([b0, b1])
	assert [int(r['streamaddress']) for r in recs] == [10, 30, 20]
	assert [bytes(arena[int(r['offset']):int(r['offset']) + int(r['len'])]) for r in recs] == [b"\x01\x02", b"\x04", b"\x03"]


def test_il2p_finish_phase_runs_rank_after_rank():
	"""IL2P decoder state is handed from shard to shard: with one worker per 'process' (threads + a barrier exchange
	standing in for the all-gather) every rank must finish only after its predecessor, with the predecessor's state."""
	import struct
	import threading
	from pymodem_b200 import _lib
	from pymodem_b200.sharded import run_protocol
	world, nc = 3, 2
	order, lock = [], threading.Lock()
	barrier = threading.Barrier(world)
	slots = [None] * world

	def make_exchange(rank):
		def exchange(blobs):
			(blob,) = blobs
			slots[rank] = blob
			barrier.wait()
			out = list(slots)
			barrier.wait()
			return out
		return exchange

	class FakeEngine:
		n_chains = nc
		has_il2p = True

	class FakeWorker:
		def __init__(self, rank):
			self.rank, self.n_chains, self.engine, self.rounds = rank, nc, FakeEngine(), 0
			self.plan = dict(tail_bits=0)
		def begin(self):
			return struct.pack("<QQIIIIq", 0, 0, 1, 1, 1, 1, 10) * nc
		def handoff(self, blobs):
			return self.begin(), False
		def gather(self, blobs):
			return b"tail%d" % self.rank
		def finish(self, tails, il2p_prev=None, il2p=False):
			assert il2p
			with lock:
				order.append(self.rank)
			prev = 0 if il2p_prev is None else struct.unpack_from("<q", il2p_prev, 0)[0]
			assert prev == self.rank * 100          # rank r sees exactly what rank r-1 produced
			st = (_lib.Il2pState * nc)()
			for c in range(nc):
				st[c].pos = (self.rank + 1) * 100
			self.il2p_blob = bytes(st)
			return struct.pack("<qq", 0, 0)

	results = [None] * world

	def run(rank):
		ex = make_exchange(rank)
		results[rank] = run_protocol([FakeWorker(rank)], ex)
	threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
	for t in threads:
		t.start()
	for t in threads:
		t.join(timeout=60)
	assert order == [0, 1, 2]
	assert all(r is not None and len(r[0]) == 0 for r in results)
