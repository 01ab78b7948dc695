"""Per-stage instruction mix of afsk_front_kernel from an `ncu --set full --import-source on` capture: the SASS page is cut
at the kernel's barriers (staging | band-pass | correlators | low-pass + epilogue) and executed instructions are
summed per opcode.  Usage: python tools/ncu_stage_mix.py gpurun_out/prof.ncu-rep > profiles/rNN_front_stage_mix.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:afsk_front"],
	capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[starts[0]]
end = starts[1] - 1 if len(starts) > 1 else len(rows)
data = [r for r in rows[starts[0] + 1:end] if len(r) == len(hdr)]
i_src, i_ex = hdr.index("Source"), hdr.index("Instructions Executed")
names = ["staging", "band-pass", "correlators", "low-pass + epilogue"]
seg = 0
per = collections.defaultdict(lambda: [0, collections.Counter()])
for r in data:
	src = r[i_src].strip()
	n = int(r[i_ex] or 0)
	op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
	per[seg][0] += n
	per[seg][1][op] += n
	if "BAR.SYNC" in src:
		seg += 1
total = sum(v[0] for v in per.values())
grid = 45000
print(f"# {rep}: first afsk_front_kernel launch of the capture, executed warp instructions per stage (cut at BAR.SYNC)")
print(f"# total {total} = {total / grid:.0f} per CTA (grid {grid})")
for k in sorted(per):
	n, ops = per[k]
	print(f"{names[k] if k < len(names) else k:22s} {100.0 * n / total:5.1f} %  {n / grid:8.0f} per CTA   " +
		"  ".join(f"{o} {100.0 * c / n:.1f}%" for o, c in ops.most_common(9)))
