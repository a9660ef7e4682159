"""Summarise .ncu-rep files (read here, no GPU): duration, DRAM bytes/throughput, occupancy,
registers, top stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_static"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "no data"); continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            name = d.get("Kernel Name", "?")[:60]
            print(f"== {path} :: {name}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:75s} {d[k]:>16s} {u.get(k, '')}")
            stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                             if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")),
                            reverse=True)[:5]
            for v, k in stalls:
                print(f"   stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")


if __name__ == "__main__":
    main()
