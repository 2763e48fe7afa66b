"""Turn an .ncu-rep (ncu --set full) into the short text summary committed under profiles/.
Usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
    print(f"# summary of {path} (ncu --set full --clock-control none; cold-cache, serialised launches)")
    for r in data:
        print("\n== " + r[idx["Kernel Name"]])
        for w in WANT:
            if w in idx:
                print(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
        vals = [(float(r[idx[h]].replace(",", "") or 0), h) for h in stalls]
        tot = sum(v for v, _ in vals) or 1
        print("  warp-state samples: " + ", ".join(f"{h[33:]} {100 * v / tot:.1f}%" for v, h in sorted(vals, reverse=True)[:8]))


if __name__ == "__main__":
    main(sys.argv[1])
