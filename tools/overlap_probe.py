"""Do the slicer (latency bound, few warps, 32 registers) and the front end (FP32 bound, register limited) overlap when
they come from different streams?  Two engines on one GPU, driven from two host threads, against one engine alone."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
lines = configs.afsk_1200_ax25_super_opt()
stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(lines)]
audio = synth.afsk1200_ax25(duration_s=1800.0, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda(); torch.cuda.synchronize()
n = len(audio)
engs = [Engine(stack), Engine(stack)]
for e in engs:
	for _ in range(3):
		e.run_device_ptr(dev.data_ptr(), n)
K = 20
t0 = time.perf_counter()
for _ in range(K):
	engs[0].run_device_ptr(dev.data_ptr(), n)
single = (time.perf_counter() - t0) / K
def work(e):
	for _ in range(K):
		e.run_device_ptr(dev.data_ptr(), n)
ths = [threading.Thread(target=work, args=(e,)) for e in engs]
t0 = time.perf_counter()
for t in ths: t.start()
for t in ths: t.join()
both = (time.perf_counter() - t0) / (2 * K)
print(f"one engine: {single * 1e3:.3f} ms per half-hour step; two engines interleaved: {both * 1e3:.3f} ms per step ({single / both:.2f}x)")
print(engs[0].stats())
