"""Executed warp-instructions per CUDA source line: python tools/ncu_inst.py <rep> [top] [launch-index]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; out = {}; hdr = None; launch = -1
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; ii = hdr.index("Instructions Executed"); continue
    if hdr and len(r) > ii and r[0] != "":
        try: n = float(r[ii])
        except Exception: continue
        key = (cur_file, r[0], r[1][:110])
        out[key] = out.get(key, 0) + n
tot = sum(out.values()) or 1
print("total warp-instructions", tot)
for (f, l, src), n in sorted(out.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{n:12.0f} {100*n/tot:5.1f}% {f}:{l} | {src}")
