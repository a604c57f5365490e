"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v_us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
    name = re.sub(r"<.*", "", r[ki])[:70]
    agg[name][0] += v_us; agg[name][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"total {tot/1e3:.2f} ms over {sum(v[1] for v in agg.values())} launches")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{t/1e3:9.3f} ms {100*t/tot:5.1f}%  n={n:5d}  {k}")
