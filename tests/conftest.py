import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
	sys.path.insert(0, REPO)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
	from oracle import oracle as orc
	orc.build()
	return orc


@pytest.fixture(scope="session")
def cuda_lib():
	"""The built CUDA library (compiled here by nvcc if missing; no GPU needed to load it)."""
	from pymodem_b200 import build, _lib
	build.build()
	return _lib.load()
