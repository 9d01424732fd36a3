"""Cut the LAST complete TTA step (gather_pack .. the launch before the next gather_pack) out of an
ncu launch list and summarise it (see summarize_launches.py)."""
import csv, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
starts = [i for i, r in enumerate(rows) if "gather_pack" in r["Kernel Name"]]
a, b = (starts[-2], starts[-1]) if len(starts) >= 2 else (starts[-1], len(rows))
tmp = dst + ".tmp.csv"
with open(tmp, "w") as f:
    w = csv.DictWriter(f, fieldnames=rows[0].keys()); w.writeheader(); w.writerows(rows[a:b])
import subprocess, os
subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "summarize_launches.py"), tmp, dst, title], check=True)
os.remove(tmp)
