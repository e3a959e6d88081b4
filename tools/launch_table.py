"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (second half = warm pass)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki = H.index("Kernel Name"); vi = H.index("Metric Value"); ui = H.index("Metric Unit")
data = rows[hdr + 1:]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else len(data) // 2
agg = collections.OrderedDict()
for r in data[skip:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("nsecond", "ns") else v / 1e3 if u in ("usecond", "us") else v
    a = agg.setdefault(r[ki][:100], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.2f} ms over {len(data) - skip} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.2f} ms {100 * v[1] / tot:5.1f}% {v[0]:5d}  {k}")
