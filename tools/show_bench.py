"""print the headline and the per-kernel-family table of a bench.py log"""
import json
import sys

for path in sys.argv[1:]:
    for l in open(path):
        l = l.strip()
        if not l.startswith("{"):
            continue
        d = json.loads(l)
        k = d.pop("kernels", None)
        print(path, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", d.get("e2e") and round(d["e2e"]["value"]),
              "launches", d.get("gpu_launches"), "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 3),
              "step_frac", round(d["step_roofline"]["frac"], 3))
        if k and "-k" in sys.argv[0:1] + sys.argv:
            for e in k:
                print(f"  {e['kernel']:45s} {e['ms_per_step']:.3f} x{e['calls_per_step']:.0f}")
