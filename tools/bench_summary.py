import json, sys
d = json.loads(sys.stdin.read())
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "clocks", d["clocks"])
r = d["roofline"]
print("roofline:", r["kernel"], r["bound"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
for k, v in r["kernels"].items():
    print(f"  {k:20s} {v['ms_per_launch']:.3f} ms x{v['launches']:<4d} share {v['share']:.3f}  alg {v['alg_tflops']:.2f} TF {v.get('alg_gbs', 0):.0f} GB/s")
