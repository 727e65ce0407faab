"""Aggregate ncu stall samples per CUDA source line: python tools/ncu_lines.py <rep.ncu-rep> [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; out = []; hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 6 and r[0] != "":
        try: samples = float(r[4])
        except Exception: continue
        out.append((samples, cur_file, r[0], r[1][:120]))
tot = sum(o[0] for o in out) or 1
for s, f, l, src in sorted(out, reverse=True)[:top]:
    print(f"{s:8.0f} {100*s/tot:5.1f}% {f}:{l} | {src}")
