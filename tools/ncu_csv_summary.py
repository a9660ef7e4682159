"""Summarise `ncu --page raw --csv` exports (tools/gpu_r2m.sh keeps these text pages instead of the
.ncu-rep files: gpurun_out/ is capped at 64 MiB) and the top stall lines of the matching
`--page source --csv --print-source sass` export.
Usage: python tools/ncu_csv_summary.py gpurun_out/<tag>/full_*.raw.csv > profiles/rN_ncu_full_summary.txt
       python tools/ncu_csv_summary.py --hot gpurun_out/<tag>/full_<kernel>.source.csv.gz [N]"""
import csv
import gzip
import io
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def summary(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            print(path, "no data"); continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            print(f"== {path.split('/')[-1]} :: {d.get('Kernel Name', '?')[:70]}")
            for k in KEYS:
                if k in d:
                    print(f"   {k:70s} {d[k]:>18s} {u.get(k, '')}")
            try:
                rd, wr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                tot = rd * scale[u["dram__bytes_read.sum"]] + wr * scale[u["dram__bytes_write.sum"]]
                t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[u["gpu__time_duration.sum"]]
                print(f"   {'=> DRAM read + write':70s} {tot / 1e6:18.1f} MB   {tot / t / 1e9:8.1f} GB/s under ncu (cold, serialised)")
            except Exception:
                pass
            stalls = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                             if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")),
                            reverse=True)[:5]
            for v, k in stalls:
                print(f"   stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.2f}")


def hot(path, n):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    si, ss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    body = [(int(r[ss] or 0), int(r[ie] or 0), r[si].strip()) for r in rows[hi + 1:] if len(r) > ss and r[0].startswith("0x")]
    tot = sum(b[0] for b in body)
    print("total samples", tot, "instructions", len(body), "executed", sum(b[1] for b in body))
    idx = sorted(range(len(body)), key=lambda i: -body[i][0])[:n]
    for i in sorted(idx):
        print("%6d %5.1f%% exec %9d  #%d  %s" % (body[i][0], 100.0 * body[i][0] / max(tot, 1), body[i][1], i, body[i][2][:90]))


if __name__ == "__main__":
    if sys.argv[1] == "--hot":
        hot(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
    else:
        summary(sys.argv[1:])
