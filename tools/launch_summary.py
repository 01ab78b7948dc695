"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file ...) -> per-kernel launches, time and share of the
LAST pass in the file (the passes repeat the same launch sequence).  usage: launch_summary.py launches.csv passes"""
import csv, sys
path, passes = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = []
with open(path) as f:
	lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
	if r.get("Metric Name") == "gpu__time_duration.sum":
		v = float(r["Metric Value"].replace(",", ""))
		unit = r.get("Metric Unit", "ns")
		ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
		rows.append((r["Kernel Name"], ms))
per = len(rows) // passes
last = rows[-per:]
acc, order = {}, []
for name, ms in last:
	name = name.split("(")[0]
	if name not in acc:
		acc[name] = [0, 0.0]; order.append(name)
	acc[name][0] += 1; acc[name][1] += ms
total = sum(v[1] for v in acc.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none, python tools/quick_run.py 3600 {passes} (last pass: {per} launches, cold-cache, serialised: compare SHARES)")
print(f"# sum of launch durations {total:.3f} ms")
for name in order:
	print(f"{name:46s} {acc[name][0]:3d} {acc[name][1]:9.4f} ms {100 * acc[name][1] / total:6.2f} %")
