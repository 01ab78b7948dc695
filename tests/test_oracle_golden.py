"""The CPU oracle against the fixtures produced by the live reference
(tools/make_golden.py) and the known-answer vectors of SURVEY.md Appendix B.2.
This is what pins the oracle; the GPU tests then compare the CUDA path to both."""
import os

import numpy as np
import pytest

from util import GOLD, Golden

CASES = ["afsk1200_superopt_48k", "afsk1200_ax25_44k1", "fsk9600_ax25_48k",
	"afsk1200_il2p_48k", "fsk9600_il2p_48k", "afsk300_real_8k"]
# recursive modems (AGC + Costas / decision-directed / PLL loops): no chunked form, AGC.apply takes
# max(buffer) of the whole recording (agc.py:67)
PSK_CASES = ["bpsk300_il2p_8k", "qpsk2400_il2p_8k", "qpsk2400_il2p_22k", "afsk300_full_8k", "bpsk1200_il2p_12k",
	"qpsk600_il2p_8k", "qpsk3600_il2p_16k", "mpsk_bpsk300_il2p_8k", "mpsk_bpsk1200_il2p_12k"]


@pytest.mark.parametrize("tag", CASES + PSK_CASES)
def test_oracle_packets_match_reference(oracle, tag):
	g = Golden(tag)
	got = oracle.run_config(g.sample_rate, g.lines, g.audio())
	assert got == g.all_packets()


@pytest.mark.parametrize("tag", CASES)
def test_oracle_chunked_equals_monolithic(oracle, tag):
	g = Golden(tag)
	got = oracle.run_config(g.sample_rate, g.lines, g.audio(), chunk=50001)
	assert got == g.all_packets()


@pytest.mark.parametrize("tag", CASES + PSK_CASES)
def test_oracle_stages_match_reference(oracle, tag):
	g = Golden(tag)
	audio = g.audio()
	for ci, line in enumerate(g.chain_lines()):
		if f"c{ci}_soft_dec" not in g.z:
			continue
		chain = oracle.Chain(g.sample_rate, line)
		soft = chain.modem.demod(audio)
		atol = 1e-9 * float(g.z[f"c{ci}_soft_rms"])
		soft_i = soft
		if isinstance(soft, tuple):                  # IQData (psk.py:748-751)
			soft_i = soft[0]
			np.testing.assert_allclose(soft[1][::97], g.z[f"c{ci}_softq_dec"], rtol=0, atol=atol)
		assert len(soft_i) == int(g.z[f"c{ci}_soft_len"])
		# same numpy.convolve calls as the reference: bit-identical on this host
		np.testing.assert_allclose(soft_i[::97], g.z[f"c{ci}_soft_dec"], rtol=0, atol=atol)
		b, a = chain.slicer.slice(soft)
		np.testing.assert_array_equal(b, g.z[f"c{ci}_sl_bytes"])
		np.testing.assert_array_equal(a, g.z[f"c{ci}_sl_addr"])
		d, _ = chain.stream.stream_unscramble_8bit(b, a)
		np.testing.assert_array_equal(d, g.z[f"c{ci}_ds_bytes"])


@pytest.mark.parametrize("tag", CASES + PSK_CASES)
def test_oracle_correlate_matches_reference(oracle, tag):
	g = Golden(tag)
	names = [l["object_name"] for l in g.chain_lines()]
	uniq, bad = oracle.correlate(g.all_packets(), names, g.sample_rate / 40)
	assert [u[0] for u in uniq] == list(g.z["uniq_addr"])
	assert [u[2] for u in uniq] == list(g.z["uniq_crc"])
	assert [len(u[3]) for u in uniq] == list(g.z["uniq_ndec"])
	assert bad == int(g.z["bad_count"])


def test_kats(oracle):
	z = np.load(os.path.join(GOLD, "kats.npz"))
	# CRC-16/X.25("123456789") = 0x906E  
	pkt = bytes(z["crc_append_123456789"])
	assert pkt[-2:] == bytes([0x6E, 0x90])
	assert oracle.crc16(b"123456789") == 0x906E
	assert oracle.check_crc(pkt) == [36974, 36974, True] == [int(x) for x in z["crc_check"]]
	data = z["lfsr_in"]
	for key in z.files:
		if not key.startswith("lfsr_") or key == "lfsr_in":
			continue
		_, poly, inv = key.split("_")
		l = oracle.LFSR({"poly": poly, "invert": "true" if inv == "1" else "false"})
		out, _ = l.stream_unscramble_8bit(data, None)
		np.testing.assert_array_equal(out, z[key])
	# SURVEY.md B.2 LFSR vectors
	l = oracle.LFSR({"poly": "0x3", "invert": "True"})
	out, _ = l.stream_unscramble_8bit(np.array([0x00, 0xFF, 0xAA, 0x0F, 0x7E], dtype=np.uint8), None)
	assert list(out) == [0xFF, 0x7F, 0x80, 0xF7, 0x3E]
	l = oracle.LFSR({"poly": "0x63003", "invert": "True"})
	out, _ = l.stream_unscramble_8bit(np.array([0x00, 0xFF, 0xAA, 0x0F, 0x7E, 0x12, 0x34], dtype=np.uint8), None)
	assert list(out) == [0xFF, 0x7F, 0x88, 0xB0, 0xF1, 0xEC, 0xA0]


def test_gf_rs_kats(oracle):
	"""GF(2^8)/0x11D tables, RS generator polynomials, RS decode."""
	table, index, inverse = oracle.gf_tables()
	assert list(table[:10]) == [1, 2, 4, 8, 16, 32, 64, 128, 29, 58]
	assert list(index[1:9]) == [0, 1, 25, 2, 50, 26, 198, 3]
	assert inverse[2] == 142 and inverse[0x53] == 140
	assert oracle.rs_genpoly(2) == [2, 3, 1]
	assert oracle.rs_genpoly(16) == [59, 36, 50, 98, 229, 41, 65, 163, 8, 30, 209, 68, 189, 104, 13, 59, 1]
	block = bytearray(32)
	block[3], block[10] = 0x55, 0xAA
	assert oracle.rs_decode(16, block) == (2, bytes(32))
	block = bytearray(32)
	for k in range(9):
		block[3 * k] = k + 1
	assert oracle.rs_decode(16, block)[0] == -1


def test_il2p_encoder_roundtrip(oracle):
	"""synth.il2p_frame (the transmitter written as the inverse of the reference's receiver) through
	the oracle's IL2P decoder: clean frames come back as the reconstructed AX.25 frames; up to 8 byte
	errors per block and 1 in the header are corrected and counted; more fail."""
	from pymodem_b200 import synth
	rng = np.random.default_rng(3)
	for plen in (0, 1, 40, 239, 240, 477, 1023):
		payload = bytes(int(x) for x in rng.integers(0, 256, plen))
		air, ax = synth.il2p_frame("MODEM", "NOISE", payload, dest_ssid=3, src_ssid=11)
		bits = synth.il2p_bits(air, preamble_bytes=3, postamble_bytes=3)
		bits = np.concatenate([bits, np.zeros((-len(bits)) % 8, dtype=np.uint8)])
		by = np.packbits(bits)
		got = oracle.IL2PCodec("x", {"sync_tol": "0"}).decode(by, np.arange(len(by), dtype=np.int64))
		assert [(g[1], g[2]) for g in got] == [(ax, 0)]
		assert oracle.check_crc(got[0][1])[2]
		# corrupt 1 header byte + 8 bytes of the first block
		bad = bytearray(by)
		first = 3 + 3                                   # preamble + sync word
		bad[first + 2] ^= 0x5A
		n_err = 1
		if plen >= 40:
			for j in range(8):
				bad[first + 15 + 2 * j] ^= 0x11 * (j + 1)
			n_err += 8
		got = oracle.IL2PCodec("x", {"sync_tol": "0"}).decode(np.frombuffer(bytes(bad), dtype=np.uint8),
			np.arange(len(by), dtype=np.int64))
		assert [(g[1], g[2]) for g in got] == [(ax, n_err)]
		if plen >= 40:
			bad[first + 15 + 17] ^= 0x77                    # a ninth error: the block fails, the packet is dropped
			got = oracle.IL2PCodec("x", {"sync_tol": "0"}).decode(np.frombuffer(bytes(bad), dtype=np.uint8),
				np.arange(len(by), dtype=np.int64))
			assert got == []


def test_ax25_quirks(oracle):
	"""Edge cases of ax25.py the GPU decoder must reproduce: abort does not clear
	the accumulated bytes; minimum length 18; frames share flags."""
	from pymodem_b200 import synth
	frame = synth.ax25_ui_frame("MODEM", "NOISE", b"hello world, this is a test")
	bits = list(synth.hdlc_bits(frame, preamble_flags=2, postamble_flags=1))
	# junk + abort (8 ones) directly before the frame's opening flag: junk bytes stay in data
	junk = [1, 0, 1, 1, 0, 0, 1, 0] * 3 + [1] * 8 + [0]
	stream = [0] * 5 + junk + bits + [0] * 11
	stream += [0] * ((-len(stream)) % 8)
	by = np.packbits(np.array(stream, dtype=np.uint8))
	addr = np.arange(1, len(by) + 1, dtype=np.int64) * 320
	pk = oracle.AX25Codec("x").decode(by, addr)
	assert len(pk) == 1 and pk[0][1] == frame
	# a too-short frame (17 bytes) is dropped
	short = list(synth.hdlc_bits(bytes(range(17)), preamble_flags=1, postamble_flags=1))
	short += [0] * ((-len(short)) % 8)
	by = np.packbits(np.array(short, dtype=np.uint8))
	assert oracle.AX25Codec("x").decode(by, np.arange(len(by), dtype=np.int64)) == []
