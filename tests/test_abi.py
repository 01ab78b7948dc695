"""The C-ABI library builds, loads without a GPU and exports every symbol that
include/pymodem_b200.h declares; the ctypes mirror matches the C structs; with
no GPU the product path fails loudly instead of falling back."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "pymodem_b200.h")


def declared_functions():
	src = open(HEADER).read()
	src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
	return sorted(set(re.findall(r"\b(pm_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported_and_bound(cuda_lib):
	from pymodem_b200 import _lib
	names = declared_functions()
	assert len(names) >= 20
	for name in names:
		assert hasattr(cuda_lib, name), f"{name} declared in the header but not exported"
		assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype"
	assert set(_lib.PROTOTYPES) == set(names)
	assert b"sm_100a" in cuda_lib.pm_version()


def test_struct_layouts_match_header(tmp_path):
	"""sizeof/offsetof of the ctypes mirrors == what gcc sees in the header."""
	from pymodem_b200 import _lib
	prog = tmp_path / "layout.c"
	prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pymodem_b200.h"\n'
		'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(pm_chain_desc), offsetof(pm_chain_desc, lpf),'
		' offsetof(pm_chain_desc, demap), offsetof(pm_chain_desc, loop), sizeof(pm_packet_rec),'
		' offsetof(pm_packet_rec, streamaddress), sizeof(pm_stats), sizeof(pm_shard_state), sizeof(pm_loop_desc),'
		' offsetof(pm_loop_desc, nco_wavetable), offsetof(pm_loop_desc, pd_table), offsetof(pm_loop_desc, hilbert_delay));return 0;}\n')
	exe = tmp_path / "layout"
	subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)], check=True)
	got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
	C, P = _lib.ChainDesc, _lib.PacketRec
	want = [ctypes.sizeof(C), C.lpf.offset, C.demap.offset, C.loop.offset, ctypes.sizeof(P), P.streamaddress.offset,
		ctypes.sizeof(_lib.Stats), ctypes.sizeof(_lib.ShardState), ctypes.sizeof(_lib.LoopDesc),
		_lib.LoopDesc.nco_wavetable.offset, _lib.LoopDesc.pd_table.offset, _lib.LoopDesc.hilbert_delay.offset]
	assert got == want


def test_no_cpu_fallback_without_gpu(cuda_lib):
	"""On a box without a GPU the engine refuses to start (PM_ERR_CUDA)."""
	import torch
	if torch.cuda.is_available():
		pytest.skip("a GPU is present")
	from pymodem_b200 import _lib
	h = ctypes.c_void_p()
	assert cuda_lib.pm_engine_create(0, ctypes.byref(h)) == _lib.PM_ERR_CUDA
	assert not h.value
	from pymodem_b200.engine import Engine, EngineError
	from pymodem_b200.modems_codecs import chain_builder
	from pymodem_b200 import configs
	stack = [chain_builder.build_chain(48000, l) for l in configs.demod_chains(configs.afsk_1200_ax25_super_opt())]
	with pytest.raises(EngineError):
		Engine(stack)


def test_product_path_does_not_import_oracle():
	"""pymodem_b200/ never references oracle/ (the oracle is test infrastructure)."""
	pkg = os.path.join(REPO, "pymodem_b200")
	for root, _dirs, files in os.walk(pkg):
		for f in files:
			if f.endswith((".py", ".cu", ".cuh", ".h")):
				text = open(os.path.join(root, f)).read()
				assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
				assert "liboracle" not in text, f


def test_command_line_exit_codes(tmp_path):
	"""python -m pymodem_b200 keeps the reference's usage and exit codes (pymodem.py:27-49) up to the point where
	a GPU is needed."""
	def run(*args):
		return subprocess.run([sys.executable, "-m", "pymodem_b200", *args], cwd=REPO, capture_output=True, text=True)
	assert run().returncode == 2
	assert run(str(tmp_path / "missing.json"), str(tmp_path / "missing.wav")).returncode == 3
	cfg = tmp_path / "c.json"
	cfg.write_text('{"object_type": "report", "object_name": "r", "options": {}}\n')
	assert run(str(cfg), str(tmp_path / "missing.wav")).returncode == 4
