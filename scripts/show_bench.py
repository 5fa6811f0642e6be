#!/usr/bin/env python
import json, sys
for f in sys.argv[1:]:
    for l in open(f):
        d = json.loads(l)
        r, h = d["roofline"], d["roofline_hbm"]
        print("%-28s %.4g wsps  %.4f ms/step  %s  %s=%.3f hbm=%.3f e2e=%.4g launches=%d cpu=%s" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], d["kernel"], r["bound"], r["frac"], h["frac"],
            d["e2e"]["value"], d["gpu_launches"], d["cpu_baseline"] and "%.3g" % d["cpu_baseline"]["value"]))
