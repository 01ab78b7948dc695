"""The north-star check at full size: the super-opt 8-chain config on the bench's one hour of 48 kHz synthetic audio --
the packet set of the CUDA engine (payload bytes, streamaddress, BytesCorrected) against the CPU oracle run chunked with
one process per chain.  Prints per-chain counts and a digest of the packet set."""
import hashlib, multiprocessing as mp, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np

SECONDS = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0


def make_audio():
	from pymodem_b200 import synth
	return synth.afsk1200_ax25(duration_s=SECONDS, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
		seed=1000, noise_seed=1001)[0]


def oracle_chain(args):
	ci, audio = args
	from oracle import oracle as orc
	from pymodem_b200 import configs
	line = configs.demod_chains(configs.afsk_1200_ax25_super_opt())[ci]
	return orc.Chain(48000, line).process_chunked(audio, chunk=1 << 20)


def digest(per_chain):
	h = hashlib.sha256()
	for ci, plist in enumerate(per_chain):
		for a, d, c in plist:
			h.update(repr((ci, int(a), bytes(d), int(c))).encode())
	return h.hexdigest()[:16]


if __name__ == "__main__":
	from oracle import oracle as orc
	orc.build()
	audio = make_audio()
	from pymodem_b200 import configs
	from pymodem_b200.engine import Engine
	from pymodem_b200.modems_codecs import chain_builder
	stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]
	eng = Engine(stack)
	t0 = time.perf_counter()
	got = [[(p.streamaddress, bytes(p.data), p.BytesCorrected) for p in pl] for pl in eng.run(audio)]
	t_gpu = time.perf_counter() - t0
	valid = sum(1 for pl in eng.run(audio) for p in pl if p.ValidCRC and p.ValidHeader)
	eng.close()
	t0 = time.perf_counter()
	with mp.get_context("fork").Pool(min(8, os.cpu_count() or 1)) as pool:
		want = pool.map(oracle_chain, [(ci, audio) for ci in range(8)])
	t_cpu = time.perf_counter() - t0
	same = got == want
	print(f"{SECONDS:g} s x 8 chains: GPU {[len(x) for x in got]} packets ({valid} with valid CRC+header) in {t_gpu * 1e3:.1f} ms incl. "
		f"Python record conversion; oracle {[len(x) for x in want]} in {t_cpu:.1f} s on {min(8, os.cpu_count() or 1)} processes")
	print("packet sets identical:", same, " digest gpu", digest(got), "oracle", digest(want))
	sys.exit(0 if same else 1)
