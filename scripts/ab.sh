#!/bin/bash
# same-box A/B of bench.py variants: scripts/ab.sh "<label>|<env assignments>|<extra bench args>" ...
# prints value / ms_per_step per variant (device-timed, 40 steps), two rounds to expose box drift
for round in 1 2; do
  for spec in "$@"; do
    IFS='|' read -r label envs extra <<< "$spec"
    out=$(env $envs python bench.py --steps 40 --warmup 5 --skip-cpu --skip-gpu-baseline --sustain-s 0 $extra 2>gpurun_out/ab_${label}.err | tail -1)
    echo "$round $label $(python -c "import json,sys; d=json.loads(sys.argv[1]); print('ms_per_step %.4f value %.1f e2e %.1f conv_ms %.3f stream_ms %.3f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['ms_per_step'], d['roofline_hbm']['ms_per_step']))" "$out" 2>/dev/null || tail -3 gpurun_out/ab_${label}.err)"
  done
done
