import json, sys
d = json.load(open(sys.argv[1]))
print("value %.4g obs/s  ms/step %.3f  pcg/step %.1f  rejects %s  rmse %.4f" % (d["value"], d["ms_per_step"], d["pcg_iters_per_step"], d["rejects"], d["final_rmse_px"]))
print("e2e %.4g obs/s  total %.3fs setup %.3fs" % (d["e2e"]["value"], d["e2e"]["seconds"], d["e2e"].get("setup_seconds", -1)))
print("roofline", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["roofline"].items()})
print("profile_pass", d.get("profile_pass"), "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in d["kernels"].items():
    print("  %-14s %8.3f ms/step  %6.1f launches/step  %9.1f us/launch  %s GB/s" % (k, v["ms_per_step"], v["launches_per_step"], v["us_per_launch"], v["achieved_gbs"] and round(v["achieved_gbs"])))
if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
