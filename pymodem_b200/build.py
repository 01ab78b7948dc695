"""Builds libpymodem_b200.so in-tree with nvcc for sm_100a (no JIT cache: the
built .so travels with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpymodem_b200.so")
SOURCES = ["engine.cu", "front.cu", "lpf_tc.cu", "slicer.cu", "bits.cu", "il2p.cu", "loops.cu", "link.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
	"-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-Xcompiler", "-pthread", "--fmad=true", "-Xptxas", "-v"]


def needs_build():
	if not os.path.exists(LIB):
		return True
	t = os.path.getmtime(LIB)
	deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pymodem_b200.h")]
	return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
	if not force and not needs_build():
		return LIB
	nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
	objs = []
	os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
	log = []
	for src in SOURCES:
		obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
		cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS
		r = subprocess.run(cmd, capture_output=True, text=True)
		log.append(r.stderr)
		if r.returncode != 0:
			sys.stderr.write(r.stdout + r.stderr)
			raise RuntimeError(f"nvcc failed on {src}")
		objs.append(obj)
	cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-pthread"]
	r = subprocess.run(cmd, capture_output=True, text=True)
	if r.returncode != 0:
		sys.stderr.write(r.stdout + r.stderr)
		raise RuntimeError("nvcc link failed")
	with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
		f.write("\n".join(log))
	if verbose:
		print("\n".join(log))
	return LIB


if __name__ == "__main__":
	build(force="--force" in sys.argv, verbose=True)
	print(LIB)
