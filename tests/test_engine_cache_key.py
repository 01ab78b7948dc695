"""The key process_chains caches its engine under (engine.stack_fingerprint): the complete state of every block, so that a
retune between two calls takes effect as it does in the reference, which reads its blocks on every call
(chain_execute.py:32-47).  Pure host logic: no GPU."""
import numpy as np

from pymodem_b200 import configs
from pymodem_b200.engine import stack_fingerprint
from pymodem_b200.modems_codecs import chain_builder


def _stack():
	return [chain_builder.build_chain(48000, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]


def test_same_parameters_same_key_whatever_the_objects():
	a, b = _stack(), _stack()
	assert stack_fingerprint(a) == stack_fingerprint(b)
	assert hash(stack_fingerprint(a)) == hash(stack_fingerprint(b))
	assert stack_fingerprint(a) != stack_fingerprint(a[:-1])


def test_retune_and_option_strings_change_the_key():
	s = _stack()
	k0 = stack_fingerprint(s)
	s[2][1].StringOptionsRetune({'space_gain': '3.0'})            # modem
	k1 = stack_fingerprint(s)
	assert k1 != k0
	s[2][1].StringOptionsRetune({'space_gain': '1.5'})            # back to the shipped value of chain 2
	assert stack_fingerprint(s) == k0
	s[5][2].StringOptionsRetune({'lock_rate': '0.9'})             # slicer
	assert stack_fingerprint(s) != k0


def test_in_place_changes_of_tap_arrays_and_attributes_change_the_key():
	s = _stack()
	k0 = stack_fingerprint(s)
	s[0][1].input_bpf[7] += 1e-12
	assert stack_fingerprint(s) != k0
	s = _stack()
	s[1][3].polynomial ^= 0x10
	assert stack_fingerprint(s) != k0
	s = _stack()
	s[0][1].output_lpf = np.array(s[0][1].output_lpf, dtype=np.float32)      # same values to 1e-8, another dtype
	assert stack_fingerprint(s) != k0
