"""The bench line contract (driver-facing JSON of bench.py), checked on the committed lines under profiles/ and on the
reference arm run here on a tiny sample.  No GPU needed."""
import glob
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
	"vs_baseline", "dtype", "data", "config", "roofline", "clocks", "e2e", "gpu_launches")


def _latest_lines():
	files = sorted(glob.glob(os.path.join(REPO, "profiles", "r01[f-z]_bench_n*.json")) +
		glob.glob(os.path.join(REPO, "profiles", "r02[a-z]_bench_n*.json")))
	assert files, "no committed bench lines"
	return [(os.path.basename(f), json.load(open(f))) for f in files]


@pytest.mark.parametrize("name,line", _latest_lines())
def test_committed_bench_lines_carry_the_contract(name, line):
	for key in REQUIRED:
		assert key in line, (name, key)
	baseline = json.load(open(os.path.join(REPO, "BASELINE.json")))
	# BASELINE.json words its metric as "demod chain-samples/sec at 1/2/4/8 B200 + decoded-packet parity vs reference":
	# the line carries the measurable part, the N and the parity are separate keys / tests
	assert baseline["metric"].startswith(line["metric"]) and line["unit"] == "chain-samples/s"
	assert line["higher_is_better"] is True and line["scaling"] == "weak" and line["data"] == "synthetic"
	assert line["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
	assert "workload" in line["config"] and "model" not in line["config"]
	assert line["warmup"] >= 3 and line["gpu_launches"] > 0
	units = line["config"]["chains"] * line["config"]["samples_per_gpu"] * line["n_gpus"]
	assert abs(line["value"] - units / (line["ms_per_step"] * 1e-3)) <= 1e-6 * line["value"]
	e2e = line["e2e"]
	# counted per rank: the rank's shard of int16 audio (plus, beyond rank 0, the slicer's warm-up history)
	assert 2 * line["config"]["samples_per_gpu"] <= e2e["h2d_bytes_per_step"] <= 2.01 * line["config"]["samples_per_gpu"]
	assert e2e["d2h_bytes_per_step"] > 0
	assert e2e["value"] < line["value"]                      # copies are inside the timed region
	roof = line["roofline"]
	for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
		assert key in roof, (name, key)
	if name.startswith("r02"):
		# round 2 lines: parity digest of what the timed steps produced, and the per-kernel rooflines at N = 1
		assert line["parity"]["match"] is True, name
		assert "frac_of_value" in e2e and "call" in e2e
		if line["n_gpus"] == 1:
			kernels = {k["kernel"]: k for k in line["roofline_all"]["kernels"]}
			assert any(k.get("share", 0) >= 0.02 and "frac" in k for k in kernels.values())
			assert e2e["h2d_ceiling"]["gbs"] > 0 and 0 < e2e["pcie_frac"] <= 1.05
	assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and 0.0 < roof["frac"] < 1.0
	clocks = line["clocks"]
	assert clocks["sm_mhz"] > 0.8 * clocks["sm_max_mhz"]
	assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
	if line["n_gpus"] == 1:
		cpu = line["cpu_baseline"]
		for key in ("value", "unit", "cores", "kind", "sample"):
			assert key in cpu, (name, key)
		assert cpu["kind"] in ("port", "reference") and cpu["unit"] == line["unit"]


def test_reference_arm_line(tmp_path):
	"""bench.py --impl reference on a two-second sample: same metric/unit/config keys, impl and cpu_baseline filled in."""
	r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
		"--cpu-seconds", "2"], capture_output=True, text=True, timeout=600, cwd=REPO)
	assert r.returncode == 0, r.stderr[-2000:]
	line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
	assert line["impl"] == "reference"
	if "unavailable" in line:
		pytest.fail("the oracle always exists: " + line["unavailable"])
	for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "config", "cpu_baseline", "e2e"):
		assert key in line, key
	assert line["cpu_baseline"]["value"] == line["value"] and line["cpu_baseline"]["kind"] in ("port", "reference")
	assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
