"""front-end tile sweep on the bench workload (one GPU): front_ms per tile size."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth, _lib
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
lines = configs.afsk_1200_ax25_super_opt()
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(lines)]
audio = synth.afsk1200_ax25(duration_s=600.0, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6,
	seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda(); torch.cuda.synchronize()
n = len(audio)
for tile in [0] + [int(x) for x in sys.argv[1:]]:
	try:
		eng = Engine(stack, tile=tile) if tile else Engine(stack)
		for _ in range(3):
			eng.run_device_ptr(dev.data_ptr(), n)
		fr = []
		for _ in range(5):
			eng.run_device_ptr(dev.data_ptr(), n); fr.append(eng.stats()['front_ms'])
		print(f"tile {tile or 'auto'} ({_lib.load().pm_engine_front_tile(eng._h, 0)}): front_ms {min(fr):.3f} (x6 = {6*min(fr):.2f} per hour)", flush=True)
		eng.close()
	except Exception as e:
		print("tile", tile, "failed:", e, flush=True)
