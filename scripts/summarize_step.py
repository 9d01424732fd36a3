"""Summarise the LAST complete TTA step of an ncu launch list that carries gpu__time_duration.sum,
dram__bytes_read.sum and dram__bytes_write.sum per launch:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:"conv_|norm_|gather_pack|head_|adam_|loss_fin|split_f32" -c 330 --csv --log-file L.csv \
        python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu
    python scripts/summarize_step.py L.csv profiles/<name>.md profiles/<traffic>.json "title"

Writes the markdown table (per-kernel totals + every launch) and a JSON with the per-step DRAM traffic
of the two kernel families bench.py reports a roofline for."""
import csv, json, sys
from collections import OrderedDict, defaultdict

src, dst, tjson, title = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else ""
rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
launches = OrderedDict()
for r in rows:
    d = launches.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "")[:40],
                                      "grid": r["Grid Size"], "block": r["Block Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"])
L = list(launches.values())
starts = [i for i, d in enumerate(L) if "gather_pack" in d["name"]]
a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (starts[-1], len(L))
step = L[a:b]
agg = defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
fam = {"conv": [0.0, 0.0], "stream": [0.0, 0.0]}
for d in step:
    t = d["gpu__time_duration.sum"] / 1e3
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    g = agg[d["name"]]
    g[0] += 1; g[1] += t; g[2] += by
    tot += t
    f = fam["conv" if "conv_" in d["name"] else "stream"]
    f[0] += t; f[1] += by
with open(dst, "w") as f:
    f.write(f"# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,"
            "dram__bytes_write.sum --clock-control none; cold-cache, serialised launches -- compare SHARES, "
            f"not absolutes)\n\nTotal {tot:.1f} us over {len(step)} launches; conv family {fam['conv'][0]:.1f} us / "
            f"{fam['conv'][1] / 1e6:.0f} MB DRAM, streaming family {fam['stream'][0]:.1f} us / "
            f"{fam['stream'][1] / 1e6:.0f} MB DRAM\n\n| kernel | launches | total us | share | DRAM MB | GB/s |\n"
            "|---|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[2] / 1e6:.1f} | {v[2] / v[1] / 1e3:.0f} |\n")
    f.write("\n## every launch (one step)\n\n| # | kernel | grid | block | us | DRAM rd MB | DRAM wr MB |\n|---|---|---|---|---|---|---|\n")
    for i, d in enumerate(step):
        f.write(f"| {i} | {d['name']} | {d['grid']} | {d['block']} | {d['gpu__time_duration.sum'] / 1e3:.1f} | "
                f"{d.get('dram__bytes_read.sum', 0) / 1e6:.1f} | {d.get('dram__bytes_write.sum', 0) / 1e6:.1f} |\n")
json.dump({"source": src, "conv_dram_bytes_per_step": fam["conv"][1], "stream_dram_bytes_per_step": fam["stream"][1],
           "conv_us": fam["conv"][0], "stream_us": fam["stream"][0], "launches": len(step)}, open(tjson, "w"), indent=1)
print(open(dst).read()[:2600])
