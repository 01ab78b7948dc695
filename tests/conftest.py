import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
	sys.path.insert(0, REPO)


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _b200_present():
	"""True when pm_engine_create would succeed: a CUDA device of compute capability 10.x.  Asked of the driver through
	torch so that collecting the tests never initialises a CUDA context in this process."""
	try:
		import torch
		return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
	except Exception:
		return False


def pytest_collection_modifyitems(config, items):
	"""A plain `pytest` on a machine without a B200 skips the gpu-marked tests (they would all fail in
	pm_engine_create -- there is no CPU fallback -- and bury real CPU-side failures).  An explicit `-m gpu` run is left
	alone: on the GPU box a missing device must fail loudly, not skip."""
	if "gpu" in (config.getoption("-m") or ""):
		return
	if _b200_present():
		return
	skip = pytest.mark.skip(reason="no sm_100 GPU: the engine has no CPU fallback (run with -m gpu on a B200)")
	for item in items:
		if "gpu" in item.keywords:
			item.add_marker(skip)


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
