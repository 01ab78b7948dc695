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
		baud=1200.0, first_frame_s=0.5, deemphasis=False, payload_len=None):
	"""int16 mono audio: AX.25 UI frames 'MODEM-0 < NOISE-0' every
	frame_interval_s, Bell-202 continuous-phase AFSK, AWGN sigma ramped linearly
	from noise_start to noise_end (in units of the signal amplitude).
	payload_len: information-field bytes per frame (an int, or a list cycled through) instead of the default 40-odd --
	frames beyond 1023 bytes exercise the max_packet_length overflow of ax25.py:46-51.
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
		if payload_len is None:
			payload = _default_payload(k, rng)
		else:
			want = payload_len[k % len(payload_len)] if isinstance(payload_len, (list, tuple)) else payload_len
			head = (f"packet {k} ").encode()
			payload = head + bytes(int(c) for c in rng.integers(0, 256, size=max(0, int(want) - len(head))))
		frame = ax25_ui_frame("MODEM", "NOISE", payload)
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


# ---------------------------------------------------------------------------------------
# IL2P transmitter (the reference only decodes, il2p.py).  Written as the inverse of the
# reference's receiver: header packing inverts unpack_il2p_header (il2p.py:214-290), the
# scrambler inverts block_unscramble (il2p.py:160-163: LFSR 0x211, receiver state 0x1F0), parity
# is the remainder modulo the RS generator (first root 0, GF(2^8)/0x11D; rs_functions.py:9-31).
# ---------------------------------------------------------------------------------------
_GF_EXP = [0] * 512
_GF_LOG = [0] * 256


def _gf_setup():
	x = 1
	for i in range(255):
		_GF_EXP[i] = x
		_GF_LOG[x] = i
		x <<= 1
		if x & 0x100:
			x ^= 0x11D
	for i in range(255, 512):
		_GF_EXP[i] = _GF_EXP[i - 255]


_gf_setup()


def _gf_mul(a, b):
	return 0 if a == 0 or b == 0 else _GF_EXP[_GF_LOG[a] + _GF_LOG[b]]


def rs_parity(data, num_roots):
	"""Systematic RS parity: remainder of data(x) * x^r modulo prod_{i<r}(x + a^i); data[0] is the
	highest-order coefficient (the receiver's Horner syndromes, rs_functions.py:36-42)."""
	gen = [1]
	for i in range(num_roots):
		nxt = [0] * (len(gen) + 1)
		for j, g in enumerate(gen):          # gen is highest-order first
			nxt[j] ^= g
			nxt[j + 1] ^= _gf_mul(g, _GF_EXP[i])
		gen = nxt
	rem = [0] * num_roots
	for d in data:
		fb = d ^ rem[0]
		rem = rem[1:] + [0]
		if fb:
			for j in range(num_roots):
				rem[j] ^= _gf_mul(gen[j + 1], fb)
	return rem


def il2p_scramble(block):
	"""Transmit side of the receiver's per-block descrambler (state 0x1F0, poly 0x211)."""
	sr = 0x1F0
	out = []
	for byte in block:
		o = 0
		for i in range(8):
			d = (byte >> (7 - i)) & 1
			s = d ^ (sr & 1)
			if s:
				sr ^= 0x211
			sr >>= 1
			o = (o << 1) | s
		out.append(o)
	return out


_HAMMING_ENCODE = [0x0, 0x71, 0x62, 0x13, 0x54, 0x25, 0x36, 0x47, 0x38, 0x49, 0x5a, 0x2b, 0x6c, 0x1d, 0x0e, 0x7f]


def il2p_frame(dest, src, payload, dest_ssid=0, src_ssid=0, command=True, trailing_crc=True):
	"""IL2P type-1 (translated AX.25 UI, PID 0xF0) frame -> (bytes on the air after the sync word,
	the AX.25 frame incl. FCS that the reference's decoder reconstructs)."""
	payload = list(payload)
	count = len(payload)
	if count > 1023:
		raise ValueError("IL2P payload is at most 1023 bytes")
	dcall = [ord(c) for c in dest.ljust(6)[:6]]
	scall = [ord(c) for c in src.ljust(6)[:6]]
	hdr = [(c - 0x20) & 0x3F for c in dcall] + [(c - 0x20) & 0x3F for c in scall] + [((dest_ssid & 0xF) << 4) | (src_ssid & 0xF)]
	hdr[0] |= 0x40                                   # UI
	hdr[1] |= 0x80                                   # header type 1
	for i in range(10):
		if count & (0x200 >> i):
			hdr[i + 2] |= 0x80
	pid_field = 0xF                                  # AX.25 PID 0xF0
	for i in range(4):
		if pid_field & (0x8 >> i):
			hdr[i + 1] |= 0x40
	control = (5 << 3) | (0x4 if command else 0)     # UI opcode 5 (control byte 0x03), C bit
	for i in range(7):
		if control & (0x40 >> i):
			hdr[i + 5] |= 0x40
	sh = il2p_scramble(hdr)
	air = sh + rs_parity(sh, 2)
	if count:
		n_blocks = -(-count // 239)
		small = count // n_blocks
		big = count - n_blocks * small
		pos = 0
		for b in range(n_blocks):
			size = small + 1 if b < big else small
			sb = il2p_scramble(payload[pos:pos + size])
			air += sb + rs_parity(sb, 16)
			pos += size
	# what the receiver rebuilds (il2p.py:292-344)
	ax = [c << 1 for c in dcall] + [((dest_ssid & 0xF) << 1) + 0x60 + (0x80 if command else 0)]
	ax += [c << 1 for c in scall] + [((src_ssid & 0xF) << 1) + 0x60 + (0 if command else 0x80) + 1]
	ax += [0x03, 0xF0] + payload
	fcs = crc16_x25(ax)
	if trailing_crc:
		air += [_HAMMING_ENCODE[(fcs >> (12 - 4 * i)) & 0xF] for i in range(4)]
	return bytes(air), bytes(ax + [fcs & 0xFF, fcs >> 8])


def il2p_bits(air, preamble_bytes=12, postamble_bytes=2):
	"""0x55 preamble + sync word 0xF15E48 + frame, MSB first -> uint8 array of bits."""
	by = [0x55] * preamble_bytes + [0xF1, 0x5E, 0x48] + list(air) + [0x55] * postamble_bytes
	return np.unpackbits(np.array(by, dtype=np.uint8))


def _add_noise_and_quantise(sig, amplitude, noise_start, noise_end, noise_seed):
	n = len(sig)
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
	return out


def _il2p_payloads(k, rng, payload_len):
	if payload_len is None:
		return _default_payload(k, rng)
	return bytes(int(c) for c in rng.integers(0, 256, size=payload_len))


def afsk1200_il2p(duration_s, sample_rate=48000, frame_interval_s=1.0, amplitude=0.5, noise_start=0.0,
		noise_end=1.0, seed=0, noise_seed=1, mark=1200.0, space=2200.0, baud=1200.0, first_frame_s=0.3,
		payload_len=None, invert=False):
	"""IL2P+CRC frames as plain (non-NRZI) Bell-202 AFSK: bit 1 = mark tone (afsk.py:162 mark - space >= 0).
	Returns (audio, [reconstructed AX.25 frames], starts)."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	sig = np.zeros(n, dtype=np.float32)
	frames, starts = [], []
	sps = sample_rate / baud
	t, k = first_frame_s, 0
	while True:
		start = int(round(t * sample_rate))
		plen = payload_len[k % len(payload_len)] if isinstance(payload_len, (list, tuple)) else payload_len
		air, ax = il2p_frame("MODEM", "NOISE", _il2p_payloads(k, rng, plen))
		line = il2p_bits(air)
		if invert:
			line = 1 - line
		nsamp = int(np.floor(len(line) * sps))
		if start + nsamp >= n:
			break
		idx = np.minimum((np.arange(nsamp) / sps).astype(np.int64), len(line) - 1)
		freq = np.where(line[idx] == 1, mark, space)
		sig[start:start + nsamp] = np.sin(2.0 * np.pi * np.cumsum(freq) / sample_rate).astype(np.float32)
		frames.append(ax)
		starts.append(start)
		k += 1
		t += max(frame_interval_s, nsamp / sample_rate + 0.1)
	return _add_noise_and_quantise(sig, amplitude, noise_start, noise_end, noise_seed), frames, starts


def fsk9600_il2p(duration_s, sample_rate=48000, frame_interval_s=0.25, amplitude=0.5, noise_start=0.0,
		noise_end=0.7, seed=0, noise_seed=1, baud=9600.0, first_frame_s=0.05, payload_len=None, trailing_crc=True):
	"""IL2P(+CRC) frames as a two-level baseband waveform (bit 1 = positive), idle = 0x55 pattern."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	sps = sample_rate / baud
	nbits = int(n / sps)
	bits = np.tile(np.array([0, 1], dtype=np.uint8), nbits // 2 + 1)[:nbits]
	frames, starts = [], []
	t, k = first_frame_s, 0
	while True:
		b0 = int(round(t * baud))
		plen = payload_len[k % len(payload_len)] if isinstance(payload_len, (list, tuple)) else payload_len
		air, ax = il2p_frame("MODEM", "NOISE", _il2p_payloads(k, rng, plen), trailing_crc=trailing_crc)
		fb = il2p_bits(air, preamble_bytes=4, postamble_bytes=1)
		if b0 + len(fb) + 64 >= nbits:
			break
		bits[b0:b0 + len(fb)] = fb
		frames.append(ax)
		starts.append(int(b0 * sps))
		k += 1
		t += max(frame_interval_s, len(fb) / baud + 0.02)
	idx = np.minimum((np.arange(n) / sps).astype(np.int64), nbits - 1)
	sig = bits[idx].astype(np.float32) * 2.0 - 1.0
	sig = np.convolve(sig, np.array([0.25, 0.5, 0.25], dtype=np.float32), 'same').astype(np.float32)
	return _add_noise_and_quantise(sig, amplitude, noise_start, noise_end, noise_seed), frames, starts


def _rrc_pulse(t, T, beta):
	"""Root-raised-cosine pulse h(t) (unit symbol energy), evaluated at arbitrary times."""
	t = np.asarray(t, dtype=np.float64)
	out = np.empty_like(t)
	x = t / T
	sing0 = np.abs(x) < 1e-9
	sing1 = np.abs(np.abs(4 * beta * x) - 1.0) < 1e-9
	reg = ~(sing0 | sing1)
	xr = x[reg]
	out[reg] = (np.sin(np.pi * xr * (1 - beta)) + 4 * beta * xr * np.cos(np.pi * xr * (1 + beta))) \
		/ (np.pi * xr * (1 - (4 * beta * xr) ** 2))
	out[sing0] = 1 - beta + 4 * beta / np.pi
	out[sing1] = beta / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta)) + (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
	return out


def _shape_symbols(symbols, n, sample_rate, baud, beta, span=6):
	"""sum_k a_k h(t - kT) sampled at n/sample_rate (any samples-per-symbol ratio); symbols may be complex."""
	T = 1.0 / baud
	sps = sample_rate / baud
	half = int(np.ceil(span * sps / 2)) + 1
	sig = np.zeros(n + 2 * half + 2, dtype=np.complex128 if np.iscomplexobj(symbols) else np.float64)
	offs = np.arange(-half, half + 1)
	K = len(symbols)
	for k0 in range(0, K, 4096):
		ks = np.arange(k0, min(K, k0 + 4096))
		centre = np.rint(ks * sps).astype(np.int64)
		idx = centre[:, None] + offs[None, :]
		tt = idx / sample_rate - (ks * T)[:, None]
		contrib = symbols[ks][:, None] * _rrc_pulse(tt, T, beta)
		ok = (idx >= 0) & (idx < n)
		np.add.at(sig, idx[ok], contrib[ok])
	return sig[:n]


def _psk_il2p_bits(duration_s, baud_bits, frame_interval_s, first_frame_s, payload_len, rng):
	"""The bit stream of a whole recording: 0x55 idle with IL2P+CRC frames dropped in."""
	nbits = int(duration_s * baud_bits)
	bits = np.tile(np.array([0, 1], dtype=np.uint8), nbits // 2 + 1)[:nbits]
	frames, starts = [], []
	t, k = first_frame_s, 0
	while True:
		b0 = int(round(t * baud_bits)) & ~7
		plen = payload_len[k % len(payload_len)] if isinstance(payload_len, (list, tuple)) else payload_len
		air, ax = il2p_frame("MODEM", "NOISE", _il2p_payloads(k, rng, plen))
		fb = il2p_bits(air, preamble_bytes=8, postamble_bytes=2)
		if b0 + len(fb) + 64 >= nbits:
			break
		bits[b0:b0 + len(fb)] = fb
		frames.append(ax)
		starts.append(b0)
		k += 1
		t += max(frame_interval_s, len(fb) / baud_bits + 0.05)
	return bits, frames, starts


def bpsk300_il2p(duration_s, sample_rate=8000, frame_interval_s=3.0, amplitude=0.5, noise_start=0.0,
		noise_end=0.6, seed=0, noise_seed=1, carrier=1500.0, baud=300.0, rolloff=0.6, first_frame_s=1.5,
		payload_len=None, carrier_phase=0.7):
	"""Differentially encoded BPSK (a 1 keeps the phase, a 0 flips it: the receive side is
	configs/bpsk_300.json's LFSR poly 0x3 + invert), RRC shaped, on a carrier that may be offset from
	the nominal 1500 Hz.  Returns (audio, [reconstructed AX.25 frames], [frame start bit])."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	bits, frames, starts = _psk_il2p_bits(duration_s, baud, frame_interval_s, first_frame_s, payload_len, rng)
	level = np.cumsum(1 - bits.astype(np.int64)) & 1          # in[n] = in[n-1] ^ (1 - d)
	base = _shape_symbols(level * 2.0 - 1.0, n, sample_rate, baud, rolloff)
	base /= max(np.max(np.abs(base)), 1e-12)
	tt = np.arange(n) / sample_rate
	sig = (base * np.cos(2.0 * np.pi * carrier * tt + carrier_phase)).astype(np.float32)
	return _add_noise_and_quantise(sig, amplitude, noise_start, noise_end, noise_seed), frames, starts


_QPSK_DEMAP = [3, 1, 2, 0, 2, 3, 0, 1, 1, 0, 3, 2, 0, 2, 1, 3]      # reference slicer.py:142-147


def qpsk2400_il2p(duration_s, sample_rate=8000, frame_interval_s=1.0, amplitude=0.5, noise_start=0.0,
		noise_end=0.5, seed=0, noise_seed=1, carrier=1500.0, baud=1200.0, rolloff=0.9, first_frame_s=1.5,
		payload_len=None, carrier_phase=0.4):
	"""Differential QPSK, 2 bits per symbol: each dibit picks the next quadrant so that the reference's
	QuadratureSlicer demap table (indexed by previous and current I/Q signs) returns it."""
	rng = np.random.default_rng(seed)
	n = int(round(duration_s * sample_rate))
	bits, frames, starts = _psk_il2p_bits(duration_s, 2 * baud, frame_interval_s, first_frame_s, payload_len, rng)
	dibits = (bits[0::2][:len(bits) // 2] << 1) | bits[1::2][:len(bits) // 2]
	nxt = np.zeros((4, 4), dtype=np.int64)
	for prev in range(4):
		for cur in range(4):
			nxt[prev, _QPSK_DEMAP[prev * 4 + cur]] = cur
	quad = np.empty(len(dibits), dtype=np.int64)
	q = 0
	for i, d in enumerate(dibits):
		q = nxt[q, d]
		quad[i] = q
	sym = (np.where(quad & 2, 1.0, -1.0) + 1j * np.where(quad & 1, 1.0, -1.0))
	base = _shape_symbols(sym, n, sample_rate, baud, rolloff)
	base /= max(np.max(np.abs(base)), 1e-12)
	tt = np.arange(n) / sample_rate
	sig = np.real(base * np.exp(1j * (2.0 * np.pi * carrier * tt + carrier_phase))).astype(np.float32)
	return _add_noise_and_quantise(sig, amplitude, noise_start, noise_end, noise_seed), frames, starts


def wav_excerpt(name):
	"""A committed excerpt of real audio (tests/golden/<name>.npz, made by tools/make_golden.py)."""
	import os
	path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name + ".npz")
	z = np.load(path)
	return np.ascontiguousarray(z["audio"], dtype=np.int16), [], []
