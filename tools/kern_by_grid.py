"""Per-(kernel, grid) duration statistics of an ncu launch list: python tools/kern_by_grid.py file.csv [skip]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]
ki, gi, vi, ui = h.index("Kernel Name"), h.index("Grid Size"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[hi + 1 + skip:]:
    try:
        v = float(r[vi].replace(",", ""))
    except (ValueError, IndexError):
        continue
    if r[ui] == "ns":
        v /= 1000
    agg[(r[ki].split("(")[0].split("::")[-1][-32:], r[gi])].append(v)
tot = sum(sum(v) for v in agg.values())
print(f"# {sys.argv[1]}: {sum(len(v) for v in agg.values())} launches, {tot:.1f} us")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v.sort()
    print(f"{k[0]:34s} {k[1]:16s} n={len(v):4d} med {v[len(v)//2]:7.2f} us  min {v[0]:7.2f}  sum {sum(v):8.1f} ({100*sum(v)/tot:4.1f}%)")
