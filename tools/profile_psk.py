"""Short runs of the float64 pipeline (qpsk_2400.json: 3 MPSK chains) and of the IL2P decoder (fsk_9600.json) for ncu."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from util import Golden
from pymodem_b200 import synth
from pymodem_b200.engine import Engine
from pymodem_b200.modems_codecs import chain_builder
for tag, audio in (("qpsk2400_il2p_8k", synth.qpsk2400_il2p(120.0, carrier=1499.0, noise_end=0.7, seed=7, noise_seed=8)[0]),
		("fsk9600_il2p_48k", synth.fsk9600_il2p(120.0, sample_rate=48000, frame_interval_s=0.25, noise_end=0.9, seed=3, noise_seed=4)[0])):
	g = Golden(tag)
	eng = Engine([chain_builder.build_chain(g.sample_rate, l) for l in g.chain_lines()])
	for _ in range(2):
		recs, _ = eng.run_raw(audio)
	print(tag, len(recs), eng.stats())
	eng.close()
