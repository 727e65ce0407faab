"""python tools/bsum.py <bench.json>: one-screen summary of a bench.py JSON line (never reads stdin)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "clocks", d["clocks"])
r = d.get("roofline") or {}
if r:
    print("roofline:", r["kernel"], r["bound"], round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), {k: r[k] for k in r if k.startswith("step_")})
    for k, v in sorted(r["kernels"].items(), key=lambda kv: -kv[1]["share"]):
        print(f"  {k:22s} {v['ms_per_launch']:.4f} ms x{v['launches']:<4d} share {v['share']:.3f}  alg {v['alg_tflops']:.2f} TF {v.get('alg_gbs', 0):.0f} GB/s")
