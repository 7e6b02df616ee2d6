import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"]), "launches", d.get("gpu_launches"))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["roofline"].items()})
tot = 0
for k in d["kernels"]:
    tot += k["ms_per_step"]
    tf = "-" if k["tflops"] is None else f"{k['tflops']:.1f}"
    print(f"{k['ms_per_step']:8.3f} ms {k['calls_per_step']:4.0f}x {tf:>7} TF {k['gbs']:7.0f} GB/s  {k['kernel']}")
print("sum", round(tot, 2))
