"""Per-kernel times of afsk_1200.json (2 AX.25 + 2 IL2P chains) on IL2P audio: where the bit level's time goes."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import torch
from pymodem_b200 import synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
from util import Golden
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 900.0
lines = Golden("afsk1200_il2p_48k").chain_lines()
audio = synth.afsk1200_il2p(sample_rate=48000, duration_s=seconds, frame_interval_s=0.8, noise_start=0.1, noise_end=1.0, seed=85,
	noise_seed=86, first_frame_s=0.3, payload_len=[None, 300, 10, 0, 240, 60])[0]
dev = torch.from_numpy(audio).cuda()
eng = Engine([chain_builder.build_chain(48000, l) for l in lines], kernel_times=1)
for _ in range(3):
	eng.run_device_ptr(dev.data_ptr(), len(audio))
print({k: round(v, 3) for k, v in eng.stats().items() if k.endswith("_ms")}, "packets", eng.stats()["n_packets"], "repairs", eng.stats()["slicer_repairs"], "segments", eng.stats()["slicer_segments"])
for name, cnt, ms in eng.kernel_times():
	print(f"{name:40s} {cnt:3d} {ms:8.3f} ms")
