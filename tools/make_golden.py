"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, importable only in the build container) on deterministic
synthetic audio.  The fixtures pin the oracle (tests/test_oracle_golden.py) and
are compared directly with the CUDA path (tests/test_gpu_*.py).

Usage: python tools/make_golden.py            (needs /root/reference)
"""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

from pymodem_b200 import synth  # noqa: E402

import modems_codecs.chain_builder as cb  # noqa: E402  (the reference's)
import modems_codecs.chain_execute as ce  # noqa: E402
import modems_codecs.crc_functions as crcf  # noqa: E402
import modems_codecs.lfsr as ref_lfsr  # noqa: E402
from modems_codecs.data_classes import AddressedData  # noqa: E402
from modems_codecs.packet_meta import PacketMetaArray  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")


def quiet():
	return contextlib.redirect_stdout(io.StringIO())


def load_config(name):
	with open(os.path.join(REF, "configs", name)) as f:
		return [json.loads(line) for line in f]


def build_ref_chain(sr, line):
	with quiet():
		m = cb.ModemConfigurator(sr, line['modem'])
		try:
			rate = m.output_sample_rate
		except AttributeError:
			rate = sr
		s = cb.SlicerConfigurator(rate, line['slicer'])
		st = cb.StreamConfigurator(line['stream'])
		c = cb.CodecConfigurator(line['codec'], line['object_name'])
	return [line['object_name'], m, s, st, c]


def run_case(tag, config_name, lines, sr, audio, meta, stage_chains=(0,)):
	"""Deterministic driver: chains in config order, fresh blocks."""
	out = {"config_json": np.array(json.dumps(lines)), "sample_rate": np.array(sr),
		"audio_sha256": np.array(hashlib.sha256(audio.tobytes()).hexdigest()),
		"meta_json": np.array(json.dumps(meta))}
	chains = [l for l in lines if l.get('object_type') == 'demod_chain']
	all_packets = []
	for ci, line in enumerate(chains):
		chain = build_ref_chain(sr, line)
		with quiet():
			soft = chain[1].demod(audio)
			sliced = chain[2].slice(soft)
			descr = chain[3].stream_unscramble_8bit(sliced)
			packets = chain[4].decode(descr)
		all_packets.append(packets)
		lens = np.array([len(p.data) for p in packets], dtype=np.int64)
		out[f"c{ci}_addr"] = np.array([p.streamaddress for p in packets], dtype=np.int64)
		out[f"c{ci}_len"] = lens
		out[f"c{ci}_corr"] = np.array([p.BytesCorrected for p in packets], dtype=np.int64)
		out[f"c{ci}_data"] = np.array([int(b) for p in packets for b in p.data], dtype=np.uint8)
		if ci in stage_chains:
			if hasattr(soft, 'i_data'):
				soft_i, soft_q = np.asarray(soft.i_data, dtype=np.float64), np.asarray(soft.q_data, dtype=np.float64)
				out[f"c{ci}_softq_dec"] = soft_q[::97]
				out[f"c{ci}_softq_win"] = soft_q[10000:10000 + 8192]
			else:
				soft_i = np.asarray(soft, dtype=np.float64)
			out[f"c{ci}_soft_len"] = np.array(len(soft_i))
			out[f"c{ci}_soft_dec"] = soft_i[::97]                 # decimated soft values
			out[f"c{ci}_soft_win"] = soft_i[10000:10000 + 8192]    # one full-rate window
			out[f"c{ci}_soft_rms"] = np.array(np.sqrt(np.mean(soft_i ** 2)))
			out[f"c{ci}_sl_bytes"] = np.array([int(d.data) for d in sliced], dtype=np.uint8)
			out[f"c{ci}_sl_addr"] = np.array([int(d.address) for d in sliced], dtype=np.int64)
			out[f"c{ci}_ds_bytes"] = np.array([int(d.data) for d in descr], dtype=np.uint8)
	# CalcCRCs + Correlate (pymodem.py:170-175)
	results = PacketMetaArray()
	for p in all_packets:
		results.add(p)
	results.CalcCRCs()
	results.Correlate(address_distance=sr / 40)
	out["uniq_addr"] = np.array([p.streamaddress for p in results.unique_packet_array], dtype=np.int64)
	out["uniq_crc"] = np.array([p.CalculatedCRC for p in results.unique_packet_array], dtype=np.int64)
	out["uniq_ndec"] = np.array([len(p.CorrelatedDecoders) for p in results.unique_packet_array], dtype=np.int64)
	out["bad_count"] = np.array(results.CountBad())
	out["n_chains"] = np.array(len(chains))
	path = os.path.join(GOLD, f"{tag}.npz")
	np.savez_compressed(path, **out)
	print(tag, "chains", len(chains), "packets", [len(p) for p in all_packets],
		"unique", len(results.unique_packet_array), "bad", results.CountBad(),
		f"{os.path.getsize(path) / 1024:.0f} KiB")


def kats():
	out = {}
	pkt = list(b"123456789")
	crcf.AppendCRC(pkt)
	out["crc_append_123456789"] = np.array(pkt, dtype=np.uint8)
	out["crc_check"] = np.array([int(x) for x in crcf.CheckCRC(pkt)], dtype=np.int64)
	rng = np.random.default_rng(7)
	data = rng.integers(0, 256, size=257, dtype=np.uint8)
	for poly, inv in ((0x3, True), (0x63003, True), (0x1, False), (0x211, False), (0x21001, False)):
		with quiet():
			l = ref_lfsr.LFSR(poly=poly, invert=inv)
			res = l.stream_unscramble_8bit([AddressedData(int(b), i) for i, b in enumerate(data)])
		out[f"lfsr_{poly:x}_{int(inv)}"] = np.array([r.data for r in res], dtype=np.uint8)
	out["lfsr_in"] = data
	np.savez_compressed(os.path.join(GOLD, "kats.npz"), **out)
	print("kats", list(out))


def main():
	os.makedirs(GOLD, exist_ok=True)
	kats()
	# (1) the headline config on 12 s of synthetic Bell-202 audio, noise ramp
	meta = dict(gen="afsk1200_ax25", duration_s=12.0, sample_rate=48000, frame_interval_s=1.0,
		noise_start=0.0, noise_end=1.4, seed=0, noise_seed=1, first_frame_s=0.3)
	audio, _, _ = synth.afsk1200_ax25(**meta_args(meta))
	run_case("afsk1200_superopt_48k", "afsk_1200_ax25_super_opt.json",
		load_config("afsk_1200_ax25_super_opt.json"), 48000, audio, meta, stage_chains=(0, 1, 7))
	# (2) same modulation at 44.1 kHz (non-integer samples/symbol: 36.75), AX.25 lines of afsk_1200.json
	meta = dict(gen="afsk1200_ax25", duration_s=8.0, sample_rate=44100, frame_interval_s=1.0,
		noise_start=0.0, noise_end=1.0, seed=2, noise_seed=3, first_frame_s=0.3)
	audio, _, _ = synth.afsk1200_ax25(**meta_args(meta))
	lines = [l for l in load_config("afsk_1200.json") if l.get('codec', {}).get('type') != 'il2p']
	run_case("afsk1200_ax25_44k1", "afsk_1200.json", lines, 44100, audio, meta, stage_chains=(0, 1))
	# (3) G3RUH FSK 9600 AX.25 (fsk_9600.json line 3: poly 0x63003 + invert)
	meta = dict(gen="fsk9600_ax25", duration_s=4.0, sample_rate=48000, frame_interval_s=0.25,
		noise_start=0.0, noise_end=0.7, seed=4, noise_seed=5, first_frame_s=0.05)
	audio, _, _ = synth.fsk9600_ax25(**meta_args(meta))
	lines = [l for l in load_config("fsk_9600.json") if l.get('codec', {}).get('type') == 'ax25'
		or l.get('object_type') == 'report']
	run_case("fsk9600_ax25_48k", "fsk_9600.json", lines, 48000, audio, meta, stage_chains=(0,))
	more_cases()
	psk_cases()
	preset_cases()


def more_cases():
	# (4) afsk_1200.json as shipped (2 AX.25 + 2 IL2P+CRC chains) on IL2P audio: RS corrections, failed
	#     headers/blocks, multi-block payloads, header-only frames
	meta = dict(gen="afsk1200_il2p", duration_s=14.0, sample_rate=48000, frame_interval_s=0.5, noise_start=0.3,
		noise_end=1.3, seed=6, noise_seed=7, first_frame_s=0.3, payload_len=[None, 300, 10, 0, 240, 60])
	audio, _, _ = synth.afsk1200_il2p(**meta_args(meta))
	run_case("afsk1200_il2p_48k", "afsk_1200.json", load_config("afsk_1200.json"), 48000, audio, meta, stage_chains=(2, 3))
	# (5) fsk_9600.json as shipped (IL2P+CRC, IL2P+CRC inverted, G3RUH AX.25) on IL2P baseband audio
	meta = dict(gen="fsk9600_il2p", duration_s=5.0, sample_rate=48000, frame_interval_s=0.25, noise_start=0.2,
		noise_end=0.9, seed=8, noise_seed=9, first_frame_s=0.05, payload_len=[None, 300, 10, 0, 240, 60])
	audio, _, _ = synth.fsk9600_il2p(**meta_args(meta))
	run_case("fsk9600_il2p_48k", "fsk_9600.json", load_config("fsk_9600.json"), 48000, audio, meta, stage_chains=(0, 1))
	# (6) the one audio file the reference ships: audio_samples/afsk_300_il2pc_noise.wav (8 kHz), first 80 s,
	#     with the correlator chains of afsk_300.json (lines 1, 4, 5; the afsk_pll lines are a different modem)
	from scipy.io.wavfile import read as readwav
	sr, wav = readwav(os.path.join(REF, "audio_samples", "afsk_300_il2pc_noise.wav"))
	wav = np.ascontiguousarray(wav[:80 * sr])
	np.savez_compressed(os.path.join(GOLD, "afsk300_wav_excerpt.npz"), audio=wav, sample_rate=np.array(sr))
	lines = [l for l in load_config("afsk_300.json") if l.get('object_type') == 'report' or l['modem']['type'] == 'afsk']
	meta = dict(gen="wav_excerpt", name="afsk300_wav_excerpt")
	run_case("afsk300_real_8k", "afsk_300.json", lines, int(sr), wav, meta, stage_chains=(0, 1))


def psk_cases():
	# (7) bpsk_300.json as shipped (Costas loop, binary slicer, differential decode, IL2P+CRC) on synthetic
	#     BPSK 300 at 8 kHz, carrier 3 Hz off nominal, noise ramp
	meta = dict(gen="bpsk300_il2p", duration_s=60.0, sample_rate=8000, frame_interval_s=3.0, noise_start=0.0,
		noise_end=0.9, seed=10, noise_seed=11, carrier=1503.0, first_frame_s=1.5, payload_len=[None, 120, 0, 30])
	audio, _, _ = synth.bpsk300_il2p(**meta_args(meta))
	run_case("bpsk300_il2p_8k", "bpsk_300.json", load_config("bpsk_300.json"), 8000, audio, meta, stage_chains=(0,))
	# (8) qpsk_2400.json as shipped (3 MPSK chains at 1475/1500/1525 Hz, quadrature slicer, IL2P+CRC) at 8 kHz
	meta = dict(gen="qpsk2400_il2p", duration_s=30.0, sample_rate=8000, frame_interval_s=1.0, noise_start=0.0,
		noise_end=0.7, seed=12, noise_seed=13, carrier=1499.0, first_frame_s=1.5, payload_len=[None, 300, 0, 30])
	audio, _, _ = synth.qpsk2400_il2p(**meta_args(meta))
	run_case("qpsk2400_il2p_8k", "qpsk_2400.json", load_config("qpsk_2400.json"), 8000, audio, meta, stage_chains=(0, 1, 2))
	# (9) the same modem at 22.05 kHz (18.375 samples per symbol; other tap counts everywhere)
	meta = dict(gen="qpsk2400_il2p", duration_s=12.0, sample_rate=22050, frame_interval_s=1.0, noise_start=0.0,
		noise_end=0.6, seed=14, noise_seed=15, carrier=1508.0, first_frame_s=1.2, payload_len=[None, 64])
	audio, _, _ = synth.qpsk2400_il2p(**meta_args(meta))
	run_case("qpsk2400_il2p_22k", "qpsk_2400.json", load_config("qpsk_2400.json")[1:], 22050, audio, meta, stage_chains=(0,))
	# (10) afsk_300.json exactly as shipped (AFSK correlator + AFSK PLL chains) on the shipped WAV excerpt
	z = np.load(os.path.join(GOLD, "afsk300_wav_excerpt.npz"))
	wav, sr = np.ascontiguousarray(z["audio"]), int(z["sample_rate"])
	meta = dict(gen="wav_excerpt", name="afsk300_wav_excerpt")
	run_case("afsk300_full_8k", "afsk_300.json", load_config("afsk_300.json"), sr, wav, meta, stage_chains=(1, 2))
	# (11) bpsk_1200.json if it is an mpsk/bpsk config: exercised through the 'bpsk' preset '1200' at 12 kHz
	meta = dict(gen="bpsk300_il2p", duration_s=12.0, sample_rate=12000, frame_interval_s=0.8, noise_start=0.0,
		noise_end=0.5, seed=16, noise_seed=17, carrier=1497.0, baud=1200.0, rolloff=0.9, first_frame_s=0.6,
		payload_len=[None, 50])
	audio, _, _ = synth.bpsk300_il2p(**meta_args(meta))
	run_case("bpsk1200_il2p_12k", "bpsk_1200.json", load_config("bpsk_1200.json"), 12000, audio, meta, stage_chains=(0,))


def _mpsk_line(name, config, carrier, lock_rate):
	"""A demod_chain line for an MPSKModem preset that no shipped config uses (psk.py:570-628: 'bpsk_300', 'bpsk_1200'),
	in the shape of configs/qpsk_2400.json: differential decoding is the stream's job (LFSR poly 0x3 + invert), the
	quadrature slicer's 'bpsk_*' presets demap the I sign (slicer.py:124-141)."""
	return {"object_name": name, "object_type": "demod_chain",
		"modem": {"type": "mpsk", "config": config, "options": {"carrier_freq": carrier}},
		"slicer": {"type": "quadrature", "config": config, "options": {"lock_rate": lock_rate}},
		"stream": {"type": "lfsr", "options": {"poly": "0x3", "invert": "True"}},
		"codec": {"type": "il2p", "options": {"crc": "yes", "disable_rs": "no", "min_dist": "0", "sync_tol": "2"}}}


def preset_cases():
	"""The presets round 1 left without a fixture (VERDICT r1 missing #3): psk.py:485-541 (qpsk_3600, qpsk_600),
	psk.py:570-628 (MPSK bpsk_300, bpsk_1200), slicer.py:124-165 (their quadrature-slicer presets)."""
	# (12) qpsk_600.json as shipped: 300 symbols/s, RRC 0.6, carrier 1500, at 8 kHz (26.67 samples per symbol)
	meta = dict(gen="qpsk2400_il2p", duration_s=50.0, sample_rate=8000, frame_interval_s=4.0, noise_start=0.0,
		noise_end=0.5, seed=20, noise_seed=21, carrier=1501.5, baud=300.0, rolloff=0.6, first_frame_s=2.0,
		payload_len=[None, 100, 0, 30])
	audio, _, _ = synth.qpsk2400_il2p(**meta_args(meta))
	run_case("qpsk600_il2p_8k", "qpsk_600.json", load_config("qpsk_600.json"), 8000, audio, meta, stage_chains=(0,))
	# (13) qpsk_3600.json as shipped: 1800 symbols/s, RRC 0.3, carrier 1650, at 16 kHz (8.89 samples per symbol)
	meta = dict(gen="qpsk2400_il2p", duration_s=14.0, sample_rate=16000, frame_interval_s=0.8, noise_start=0.0,
		noise_end=0.4, seed=22, noise_seed=23, carrier=1648.0, baud=1800.0, rolloff=0.3, first_frame_s=1.0,
		payload_len=[None, 200, 0, 40])
	audio, _, _ = synth.qpsk2400_il2p(**meta_args(meta))
	run_case("qpsk3600_il2p_16k", "qpsk_3600.json", load_config("qpsk_3600.json"), 16000, audio, meta, stage_chains=(0,))
	# (14) MPSKModem preset 'bpsk_300' (decision-directed loop on a two-point constellation) at 8 kHz
	meta = dict(gen="bpsk300_il2p", duration_s=50.0, sample_rate=8000, frame_interval_s=3.0, noise_start=0.0,
		noise_end=0.7, seed=24, noise_seed=25, carrier=1502.0, first_frame_s=1.5, payload_len=[None, 60, 0, 30])
	audio, _, _ = synth.bpsk300_il2p(**meta_args(meta))
	lines = [_mpsk_line("MPSK2 300 IL2P+CRC 1500", "bpsk_300", "1500", "0.815")]
	run_case("mpsk_bpsk300_il2p_8k", None, lines, 8000, audio, meta, stage_chains=(0,))
	# (15) MPSKModem preset 'bpsk_1200' at 12 kHz
	meta = dict(gen="bpsk300_il2p", duration_s=14.0, sample_rate=12000, frame_interval_s=0.8, noise_start=0.0,
		noise_end=0.5, seed=26, noise_seed=27, carrier=1497.0, baud=1200.0, rolloff=0.9, first_frame_s=0.6,
		payload_len=[None, 50])
	audio, _, _ = synth.bpsk300_il2p(**meta_args(meta))
	lines = [_mpsk_line("MPSK2 1200 IL2P+CRC 1500", "bpsk_1200", "1500", "0.9")]
	run_case("mpsk_bpsk1200_il2p_12k", None, lines, 12000, audio, meta, stage_chains=(0,))


def meta_args(meta):
	return {k: v for k, v in meta.items() if k != "gen"}


if __name__ == "__main__":
	if "--more" in sys.argv:
		more_cases()
	elif "--psk" in sys.argv:
		psk_cases()
	elif "--presets" in sys.argv:
		preset_cases()
	else:
		main()
