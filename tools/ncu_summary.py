"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
	"gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
	"launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
	"dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
	"sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
	"sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
	"sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
	"l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
	"smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
	"smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main(path):
	raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
	rows = list(csv.reader(io.StringIO(raw)))
	hdr, units = rows[0], rows[1]
	ki = hdr.index("Kernel Name")
	print(f"# {path}: ncu --set full --clock-control none (per launch; cold-cache, serialised)")
	for r in rows[2:]:
		print(f"\n== {r[ki].split('(')[0]}  [id {r[0]}]")
		for w in WANT:
			if w in hdr:
				i = hdr.index(w)
				print(f"  {w:80s} {r[i]:>18s} {units[i]}")


if __name__ == "__main__":
	main(sys.argv[1])
