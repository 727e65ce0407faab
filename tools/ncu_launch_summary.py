"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/ncu_launch_summary.py launches.csv"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = {}
order = []
for r in rows:
    name = re.sub(r"^void\s+", "", r[4])
    name = re.sub(r"\(.*$", "", name).replace("ccsd::", "")
    ns = float(r[14].replace(",", ""))
    if name not in agg:
        agg[name] = [0, 0.0]
        order.append(name)
    agg[name][0] += 1
    agg[name][1] += ns
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | mean ms | share |\n|---|---|---|---|---|")
for n in order:
    c, t = agg[n]
    print(f"| `{n}` | {c} | {t/1e6:.3f} | {t/1e6/c:.3f} | {t/tot:.3f} |")
