"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into profiles/<name>.md"""
import collections
import csv
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hdr]
rows = rows[hdr + 1:]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    n = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(n, []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(dst, "w") as f:
    f.write(f"# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised "
            "launches: compare SHARES, not absolutes).\n\n| kernel | launches | mean us | share of GPU time |\n|---|---|---|---|\n")
    for n, v in agg.items():
        f.write(f"| `{n[:70]}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot:.3f} |\n")
print(open(dst).read())
