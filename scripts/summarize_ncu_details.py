"""profiles/<name>.md from `ncu -i X.ncu-rep --page details --csv` of one TTA step
(sections SpeedOfLight, MemoryWorkloadAnalysis, Occupancy, LaunchStats, ComputeWorkloadAnalysis,
WarpStateStats): one row per launch with the numbers the roofline discussion in DESIGN.md uses."""
import csv, sys
from collections import OrderedDict

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
h = rows[0]
I = {k: h.index(k) for k in ("ID", "Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value", "Grid Size")}
want = OrderedDict([("Duration", "us"), ("DRAM Throughput", "dram %"), ("Memory Throughput", None),
                    ("L2 Cache Throughput", "L2 %"), ("Compute (SM) Throughput", "SM %"),
                    ("Executed Ipc Active", "IPC"), ("Registers Per Thread", "regs"),
                    ("Achieved Occupancy", "occ %"), ("L2 Hit Rate", "L2 hit %"),
                    ("Warp Cycles Per Issued Instruction", "cyc/inst")])
K = OrderedDict()
for r in rows[1:]:
    d = K.setdefault(r[I["ID"]], {"name": r[I["Kernel Name"]].split("(")[0].replace("void ", "")[:36], "grid": r[I["Grid Size"]]})
    m, v, u = r[I["Metric Name"]], r[I["Metric Value"]], r[I["Metric Unit"]]
    if m == "Memory Throughput" and "byte" in u:
        d["mem"] = f"{v} {u.replace('byte/s', 'B/s')}"
    elif m in want and m != "Memory Throughput" and m not in d:
        d[m] = v
with open(dst, "w") as f:
    f.write(f"# {title}\n\nSource: `ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy "
            "--section LaunchStats --section ComputeWorkloadAnalysis --section WarpStateStats --clock-control none` "
            "over the 104 launches of ONE step (`python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu`), "
            "exported with `--page details --csv`. Cold cache, serialised: compare shares.\n\n")
    cols = [k for k in want if k != "Memory Throughput"]
    f.write("| # | kernel | grid | " + " | ".join(want[c] for c in cols) + " | mem throughput |\n|---|---|---|" + "---|" * (len(cols) + 1) + "\n")
    for i, d in enumerate(K.values()):
        f.write(f"| {i} | {d['name']} | {d['grid']} | " + " | ".join(str(d.get(c, "")) for c in cols) + f" | {d.get('mem', '')} |\n")
print(open(dst).read()[:3000])
