python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E|passed|failed|FAILED" | head -12
python bench.py --skip-cpu 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'conv', d['roofline']['ms_per_step'], 'hbm', d['roofline_hbm']['ms_per_step'])"
