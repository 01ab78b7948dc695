"""Where the tensor-core low-pass kernel's time goes (option stage_clocks): cycles the control thread waits for data,
for a free operand slot, for the epilogue, and spends issuing; cycles the epilogue waits and works.  Per item / tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
eng = Engine([chain_builder.build_chain(48000, l) for l in lines], kernel_times=1, stage_clocks=1, tc_debug=int(sys.argv[2]) if len(sys.argv) > 2 else 0, tc_grid=int(sys.argv[3]) if len(sys.argv) > 3 else 148)
for _ in range(3):
	eng.run_device_ptr(dev.data_ptr(), len(audio))
eng.stage_clocks()
eng.run_device_ptr(dev.data_ptr(), len(audio))
c = eng.stage_clocks()
tiles = (len(audio) - 305 + 8191) // 8192
print(f"control loop: {c[7] / (tiles * 4):.0f} cycles per item ({c[7] / (tiles * 4 * 66):.0f} per MMA); issue block {c[3] / (tiles * 4):.0f}; epilogue per tile: wait {c[5] / tiles:.0f}, work {c[6] / tiles:.0f}")
print("mask", sys.argv[2] if len(sys.argv) > 2 else 0, "kernel times:", [(n, round(ms, 3)) for n, k, ms in eng.kernel_times() if "front" in n or "lpf" in n])
