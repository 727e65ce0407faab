"""Print the hottest SASS/source lines of an ncu --page source --csv dump (stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, ist = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
il = hdr.index("stall_long_sb"); ib = hdr.index("stall_barrier"); iw = hdr.index("stall_wait"); ish = hdr.index("stall_short_sb"); imio = hdr.index("stall_mio")
data = []
for n, r in enumerate(rows[2:]):
    try:
        v = float(r[ist])
    except Exception:
        continue
    data.append((v, n, r[isrc][:100], r[il], r[ib], r[ish], r[imio]))
tot = sum(d[0] for d in data) or 1
for v, n, s, l, b, sh, mio in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{v:8.0f} {100*v/tot:5.1f}%  line {n:5d} long_sb={l} bar={b} short={sh} mio={mio} | {s}")
