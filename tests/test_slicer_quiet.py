"""CPU check of the arithmetic fact the GPU slicer's repairs rely on in stretches without zero crossings
(csrc/slicer.cu, SlicerChain::quiet_words): the reference's clock loop is exactly periodic after at most one period
when samples per symbol is a dyadic rational -- and is NOT when it is not, which is why the engine only uses it there."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import slicer_quiet_check as qc


@pytest.mark.parametrize("sps", [40.0, 5.0, 10.0, 160.0, 32.0, 36.75])
def test_clock_is_exactly_periodic_without_crossings(sps):
	P, fails, worst = qc.check(sps, trials=1500)
	assert fails == 0
	assert worst < P          # the engine steps ceil(P / 32) + 1 whole words exactly before it trusts the repetition


def test_not_periodic_for_non_dyadic_rates():
	P, fails, worst = qc.check(8000.0 / 300.0, trials=50)
	assert fails == 50
