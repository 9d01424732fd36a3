#!/bin/bash
# multi-GPU measurement suite for N = $1 GPUs of one box: parity worker, headline (cfg 2, weak scaling),
# cfg 4 (18 windows) and its 48-window variant (strong scaling over windows)
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -2
$TR --master-port 29750 bench.py --gpus $N --steps 20 --warmup 5 --skip-cpu --skip-gpu-baseline > gpurun_out/mg${N}_cfg2.json 2> gpurun_out/mg${N}_cfg2.err
$TR --master-port 29751 bench.py --gpus $N --steps 6 --warmup 3 --workload cfg4 --sw-batch 1 > gpurun_out/mg${N}_cfg4.json 2> gpurun_out/mg${N}_cfg4.err
$TR --master-port 29752 bench.py --gpus $N --steps 6 --warmup 3 --workload cfg4_96 --sw-batch 2 > gpurun_out/mg${N}_cfg4_96.json 2> gpurun_out/mg${N}_cfg4_96.err
for f in cfg2 cfg4 cfg4_96; do
  python - "$f" "$N" <<'PY'
import json,sys
f,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads(open(f"gpurun_out/mg{n}_{f}.json").read().strip().splitlines()[-1])
    keys=("value","ms_per_step","windows_per_s","accumulator_allreduce_ms")
    print(f, n, {k:d.get(k) for k in keys if k in d}, "e2e", (d.get("e2e") or {}).get("value"), "sustained", (d.get("sustained") or {}).get("value"), "numa", (d.get("e2e") or {}).get("numa"), (d.get("config") or {}).get("ideal_speedup"))
except Exception as e:
    print(f, n, "FAILED", e); print(open(f"gpurun_out/mg{n}_{f}.err").read()[-1500:])
PY
done
