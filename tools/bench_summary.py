import json, sys
d = json.loads(sys.stdin.read())
print("ms/step", round(d["ms_per_step"], 3), "value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"], "clocks", d["clocks"])
for k, v in d["roofline"]["kernels"].items():
    print(f"  {k:16s} {v['ms_per_launch']:.3f} ms x{v['launches']}  share {v['share']:.3f}  alg {v['alg_tflops']:.2f} TF")
