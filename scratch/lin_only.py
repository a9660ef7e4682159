import json, sys
d = json.load(open(sys.argv[1]))
k = d["kernels"]["linearize"]
print(sys.argv[1], "linearize %.1f us" % k["us_per_launch"])
