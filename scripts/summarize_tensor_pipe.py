"""Per-launch tensor-pipe utilisation of the conv kernels of ONE TTA step, from

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,gpc__cycles_elapsed.max.per_second,\
l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:conv_ \
        --csv --log-file L.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-cpu
    python scripts/summarize_tensor_pipe.py L.csv profiles/<name>.md profiles/tensor_pipe_r1.json "title" [launches/step]
"""
import csv, json, sys
from collections import OrderedDict

src, dst, tjson, title = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = list(csv.DictReader([l for l in open(src) if not l.startswith("==")]))
L = OrderedDict()
for r in rows:
    d = L.setdefault(r["ID"], {"name": r["Kernel Name"].split("(")[0].replace("void ", "").replace("tta::", "")[:34],
                               "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
L = list(L.values())
n = int(sys.argv[5]) if len(sys.argv) > 5 else 35   # conv launches of one step (BraTS res-unit UNet)
step = L[-n:]
T, TP, TC, SM, GHZ = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                      "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
                      "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
                      "gpc__cycles_elapsed.max.per_second")
tot = sum(d[T] for d in step)
wt = sum(d[T] * d[TP] for d in step) / tot
wtc = sum(d[T] * d[TC] for d in step) / tot
ghz = sum(d[T] * d[GHZ] for d in step) / tot / 1e9
out = [f"# {title}", "",
       f"Source: `{src}` (ncu, metrics in scripts/summarize_tensor_pipe.py; last {n} conv launches = one step, forward then "
       "backward; cold cache, serialised).", "",
       f"Conv launches total {tot / 1e3:.1f} us; duration-weighted `sm__pipe_tensor_cycles_active` = {wt:.1f} % of elapsed "
       f"cycles, `sm__pipe_tc_cycles_active` = {wtc:.1f} %, SM clock {ghz:.2f} GHz.", "",
       "| # | kernel | grid | us | tensor pipe active % | tc pipe active % | smem operand wavefronts % | SM GHz |",
       "|---|---|---|---|---|---|---|---|"]
for i, d in enumerate(step):
    out.append(f"| {i} | {d['name']} | {d['grid']} | {d[T] / 1e3:.1f} | {d[TP]:.1f} | {d[TC]:.1f} | {d[SM]:.1f} | "
               f"{d[GHZ] / 1e9:.2f} |")
open(dst, "w").write("\n".join(out) + "\n")
json.dump({"source": src, "conv_us": tot / 1e3, "tensor_pipe_active_pct": wt, "tc_pipe_active_pct": wtc,
           "sm_ghz_under_load": ghz, "launches": n}, open(tjson, "w"), indent=1)
print("\n".join(out[:8]))
