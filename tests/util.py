"""Shared helpers for the parity tests."""
import hashlib
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Golden:
	"""One tests/golden/<tag>.npz made by tools/make_golden.py from the live reference."""

	def __init__(self, tag):
		self.z = np.load(os.path.join(GOLD, tag + ".npz"))
		self.lines = json.loads(str(self.z["config_json"]))
		self.sample_rate = int(self.z["sample_rate"])
		self.meta = json.loads(str(self.z["meta_json"]))
		self.n_chains = int(self.z["n_chains"])

	def audio(self):
		"""Regenerate the audio from its seeded recipe and check it is the one the
		reference saw."""
		from pymodem_b200 import synth
		args = {k: v for k, v in self.meta.items() if k != "gen"}
		audio = getattr(synth, self.meta["gen"])(**args)[0]
		assert hashlib.sha256(audio.tobytes()).hexdigest() == str(self.z["audio_sha256"]), \
			"synthetic generator drifted from the committed golden fixture"
		return audio

	def packets(self, ci):
		"""[(streamaddress, data bytes, BytesCorrected)] of chain ci"""
		addr, lens, corr = self.z[f"c{ci}_addr"], self.z[f"c{ci}_len"], self.z[f"c{ci}_corr"]
		data = self.z[f"c{ci}_data"].tobytes()
		out, off = [], 0
		for a, n, c in zip(addr, lens, corr):
			out.append((int(a), data[off:off + int(n)], int(c)))
			off += int(n)
		return out

	def all_packets(self):
		return [self.packets(ci) for ci in range(self.n_chains)]

	def chain_lines(self):
		return [l for l in self.lines if l.get("object_type") == "demod_chain"]


def as_tuples(per_chain_packetmeta):
	"""list[list[PacketMeta]] -> the oracle/golden tuple form"""
	return [[(p.streamaddress, bytes(p.data), p.BytesCorrected) for p in plist] for plist in per_chain_packetmeta]
