"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a markdown table (profiles/*.md).
    python scripts/summarize_ncu_full.py raw.csv out.md "title" [extra-metric-substring ...]"""
import csv
import sys

src, dst, title = sys.argv[1:4]
extra = sys.argv[4:]
rows = list(csv.reader(open(src)))
h, u, data = rows[0], rows[1], rows[2:]
cols = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")]
for e in extra:
    for name in h:
        if e in name and all(name != c[0] for c in cols):
            cols.append((name, name))
cols = [(c, n) for c, n in cols if c in h]
ki = h.index("Kernel Name")
with open(dst, "w") as f:
    f.write(f"# {title}\n\nSource: `{src}` exported from the `.ncu-rep` (ncu --set full --clock-control none).\n\n")
    f.write("| kernel | " + " | ".join(n for _, n in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    for r in data:
        vals = []
        for c, _ in cols:
            i = h.index(c)
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            vals.append(f"{v} {u[i]}".strip())
        f.write(f"| `{r[ki].split('(')[0][:48]}` | " + " | ".join(vals) + " |\n")
print(open(dst).read())
