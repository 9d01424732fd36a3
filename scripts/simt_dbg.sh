for v in "" "--set fuse_small_norm=false" "--set small_norm_max_voxels=4096"; do echo "== $v"; python bench.py --skip-cpu $v 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'conv', d['roofline']['ms_per_step'], 'hbm', d['roofline_hbm']['ms_per_step'])"; done
