"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smsp__inst_executed.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), "| grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"))
        for k in KEYS:
            if k in d:
                print("  %-82s %s %s" % (k, d[k], rows[1][hdr.index(k)]))
        stalls = sorted(((float(v.replace(",", "")), k[len(STALL):].replace("_per_issue_active.ratio", ""))
                         for k, v in d.items() if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v),
                        reverse=True)[:6]
        print("  top stalls (warps per issue-active cycle):", ", ".join("%s %.2f" % (n, v) for v, n in stalls))
        tr = float(d["dram__bytes_read.sum"].replace(",", "")) + float(d["dram__bytes_write.sum"].replace(",", ""))
        print("  dram traffic (read+write, unit of the columns above):", tr)


if __name__ == "__main__":
    main(sys.argv[1])
