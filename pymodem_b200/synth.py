"""Synthetic packet-audio generators (the reference ships no modulator).

Used by tests and bench.py to make the workloads BASELINE.json names:
Bell-202 AFSK 1200 AX.25 UI frames in AWGN whose sigma ramps up across the
recording (reference README.md:43-44 describes its sample files that way).
Deterministic: numpy.random.default_rng(seed) for payloads and noise.
"""
import numpy as np


def crc16_x25(data):
	"""CRC-16/X.25 (poly 0x8408 reflected, init/xorout 0xFFFF) -- the AX.25 FCS
	(same arithmetic as reference crc_functions.py:63-76)."""
	crc = 0xFFFF
	for byte in data:
		crc ^= int(byte)
		for _ in range(8):
			crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
	return crc ^ 0xFFFF


def ax25_ui_frame(dest, src, payload, dest_ssid=0, src_ssid=0):
	"""AX.25 UI frame bytes incl. FCS (low byte first)."""
	def addr(call, ssid, last):
		call = call.ljust(6)[:6]
		out = [ord(c) << 1 for c in call]
		out.append(0x60 | ((ssid & 0xF) << 1) | (1 if last else 0))
		return out
	frame = addr(dest, dest_ssid, False) + addr(src, src_ssid, True) + [0x03, 0xF0] + list(payload)
	fcs = crc16_x25(frame)
	return bytes(frame + [fcs & 0xFF, fcs >> 8])


def hdlc_bits(frame, preamble_flags=32, postamble_flags=2):
	"""Flags + bit-stuffed frame, LSB-first per byte -> uint8 array of line bits
	(before NRZI)."""
	flag = [0, 1, 1, 1, 1, 1, 1, 0]
	bits = flag * preamble_flags
	ones = 0
	for byte in frame:
		for i in range(8):
			b = (byte >> i) & 1
			bits.append(b)
			if b:
				ones += 1
				if ones == 5:
					bits.append(0)
					ones = 0
			else:
				ones = 0
	bits += flag * postamble_flags
	return np.array(bits, dtype=np.uint8)


def nrzi(bits, level=1):
	"""NRZI: a 0 toggles the line, a 1 keeps it."""
	out = np.empty(len(bits), dtype=np.uint8)
	for i, b in enumerate(bits):
		if not b:
			level ^= 1
		out[i] = level
	return out


def g3ruh_scramble(bits):
	"""G3RUH scrambler 1 + x^12 + x^17 (applied after NRZI on the TX side)."""
	sr = 0
	out = np.empty(len(bits), dtype=np.uint8)
	for i, b in enumerate(bits):
		o = int(b) ^ ((sr >> 11) & 1) ^ ((sr >> 16) & 1)
		sr = ((sr << 1) | o) & 0x1FFFF
		out[i] = o
	return out


def _default_payload(k, rng):
	chars = rng.integers(33, 127, size=40)
	return (f"packet {k} ").encode() + bytes(int(c) for c in chars)


def afsk1200_ax25(duration_s, sample_rate=48000, frame_interval_s=3.1, amplitude=0.5,
		noise_start=0.0, noise_end=1.6, seed=0, noise_seed=1, mark=1200.0, space=2200.0,
		baud=1200.0, first_frame_s=0.5, deemphasis=False):
	"""int16 mono audio: AX.25 UI frames 'MODEM-0 < NOISE-0' every
	frame_interval_s, Bell-202 continuous-phase AFSK, AWGN sigma ramped linearly
	from noise_start to noise_end (in units of the signal amplitude).
	Returns (audio int16[N], frames list of bytes, frame start samples)."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	sig = np.zeros(n, dtype=np.float32)
	frames, starts = [], []
	t = first_frame_s
	k = 0
	sps = sample_rate / baud
	while True:
		start = int(round(t * sample_rate))
		frame = ax25_ui_frame("MODEM", "NOISE", _default_payload(k, rng))
		line = nrzi(hdlc_bits(frame))
		nsamp = int(np.floor(len(line) * sps))
		if start + nsamp >= n:
			break
		idx = np.minimum((np.arange(nsamp) / sps).astype(np.int64), len(line) - 1)
		freq = np.where(line[idx] == 1, mark, space)
		phase = 2.0 * np.pi * np.cumsum(freq) / sample_rate
		tone = np.sin(phase)
		if deemphasis:
			tone = np.where(line[idx] == 1, tone, tone * (mark / space))
		sig[start:start + nsamp] = tone.astype(np.float32)
		frames.append(frame)
		starts.append(start)
		k += 1
		t += frame_interval_s
	out = np.empty(n, dtype=np.int16)
	nrng = np.random.default_rng(noise_seed)
	chunk = 1 << 22
	fs = 32767.0 * amplitude
	for pos in range(0, n, chunk):
		m = min(chunk, n - pos)
		ramp = noise_start + (noise_end - noise_start) * (np.arange(pos, pos + m, dtype=np.float32) / max(n - 1, 1))
		x = sig[pos:pos + m] + ramp * nrng.standard_normal(m, dtype=np.float32)
		np.clip(x * fs, -32768, 32767, out=x)
		out[pos:pos + m] = np.rint(x).astype(np.int16)
	return out, frames, starts


def fsk9600_ax25(duration_s, sample_rate=48000, frame_interval_s=0.5, amplitude=0.5,
		noise_start=0.0, noise_end=0.8, seed=0, noise_seed=1, baud=9600.0, first_frame_s=0.1):
	"""int16 mono audio: G3RUH-scrambled NRZI AX.25 at 9600 bd as a two-level
	baseband waveform (one-pole-ish smoothing by a short raised-cosine edge)."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	sps = sample_rate / baud
	nbits = int(n / sps)
	# Continuous scrambled stream: idle = flags, frames inserted at intervals.
	flag = np.array([0, 1, 1, 1, 1, 1, 1, 0], dtype=np.uint8)
	bits = np.tile(flag, nbits // 8 + 1)[:nbits]
	frames, starts = [], []
	t = first_frame_s
	k = 0
	while True:
		b0 = int(round(t * baud / 8)) * 8
		frame = ax25_ui_frame("MODEM", "NOISE", _default_payload(k, rng))
		fb = hdlc_bits(frame, preamble_flags=1, postamble_flags=1)
		pad = (-len(fb)) % 8
		if b0 + len(fb) + pad + 64 >= nbits:
			break
		bits[b0:b0 + len(fb)] = fb
		if pad:
			# keep the idle flag pattern byte-aligned after the frame
			bits[b0 + len(fb):b0 + len(fb) + pad] = 1
		frames.append(frame)
		starts.append(int(b0 * sps))
		k += 1
		t += frame_interval_s
	line = g3ruh_scramble(nrzi(bits))
	idx = np.minimum((np.arange(n) / sps).astype(np.int64), nbits - 1)
	sig = (line[idx].astype(np.float32) * 2.0 - 1.0)
	# soften the edges a little (a TX low-pass)
	k3 = np.array([0.25, 0.5, 0.25], dtype=np.float32)
	sig = np.convolve(sig, k3, 'same').astype(np.float32)
	nrng = np.random.default_rng(noise_seed)
	ramp = noise_start + (noise_end - noise_start) * (np.arange(n, dtype=np.float32) / max(n - 1, 1))
	x = sig + ramp * nrng.standard_normal(n, dtype=np.float32)
	x = np.clip(x * (32767.0 * amplitude), -32768, 32767)
	return np.rint(x).astype(np.int16), frames, starts
