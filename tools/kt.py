"""print per-kernel times of a bench JSON log"""
import json, sys
for p in sys.argv[1:]:
    try:
        d = json.loads(open(p).read().strip().splitlines()[-1])
        print(p, "Gpix/s %.3f ms/step %.3f step_frac %.3f" % (d["value"], d["ms_per_step"], d["roofline_step"]["frac"]),
              {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in d["kernels"].items()})
    except Exception as e:
        print(p, "unreadable:", e)
