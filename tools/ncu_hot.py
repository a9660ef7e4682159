"""Top stall-sample SASS lines of a .ncu-rep (first kernel in the report).  Usage: ncu_hot.py rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
si, ss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = []
for r in rows[hdr_i + 1:]:
    if len(r) <= ss or r[0] == "Address" or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name" and body: break
        continue
    body.append((int(r[ss] or 0), int(r[ie] or 0), r[si].strip()))
tot = sum(b[0] for b in body)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total samples", tot, "instructions", len(body), "executed", sum(b[1] for b in body))
idx = sorted(range(len(body)), key=lambda i: -body[i][0])[:n]
for i in sorted(idx):
    print("%5d %5.1f%% exec %8d  #%d  %s" % (body[i][0], 100.0 * body[i][0] / max(tot, 1), body[i][1], i, body[i][2][:90]))
