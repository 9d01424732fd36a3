"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel summary."""
import csv, sys
from collections import defaultdict

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
with open(src) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
agg, cnt, tot = defaultdict(float), defaultdict(int), 0.0
lines = []
for i, r in enumerate(rows):
    name = r["Kernel Name"].split("(")[0].replace("void ", "")[:48]
    t = float(r["Metric Value"]) / 1e3
    agg[name] += t; cnt[name] += 1; tot += t
    lines.append(f"| {i} | {name} | {r['Grid Size']} | {r['Block Size']} | {t:.1f} |")
with open(dst, "w") as f:
    f.write(f"# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, "
            f"serialised launches -- compare SHARES, not absolutes)\n\nTotal {tot:.1f} us over {len(rows)} launches\n\n"
            "| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1]):
        f.write(f"| {k} | {cnt[k]} | {v:.1f} | {100 * v / tot:.1f}% |\n")
    f.write("\n## every launch (one step)\n\n| # | kernel | grid | block | us |\n|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
print(open(dst).read()[:1500])
