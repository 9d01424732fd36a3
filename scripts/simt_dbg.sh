for v in "TTA_PDL=0" "TTA_PDL_OFF=0" "TTA_PDL_OFF=8" "TTA_PDL_OFF=2" "TTA_PDL_OFF=22" "TTA_PDL_OFF=10"; do echo "== $v"; env $v python bench.py --skip-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'conv', d['roofline']['ms_per_step'], 'hbm', d['roofline_hbm']['ms_per_step'])"; done
