"""A few device-resident passes of the bench workload and nothing else (the command ncu wraps): prints the stage times
of the last pass.  usage: quick_run.py [seconds] [passes] [key=value engine options ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pymodem_b200 import configs, synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
opts = dict(kv.split("=") for kv in sys.argv[3:])
lines = configs.demod_chains(configs.afsk_1200_ax25_super_opt())
audio = synth.afsk1200_ax25(duration_s=seconds, sample_rate=48000, frame_interval_s=3.1, noise_start=0.0, noise_end=1.6, seed=1000, noise_seed=1001)[0]
dev = torch.from_numpy(audio).cuda()
eng = Engine([chain_builder.build_chain(48000, l) for l in lines], **opts)
for _ in range(passes):
	eng.run_device_ptr(dev.data_ptr(), len(audio))
st = eng.stats()
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()})
eng.close()
