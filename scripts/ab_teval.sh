#!/bin/bash
# A/B of the t_eval mode (dense output at 1000 points, README.md:54) for variant libraries: bench.py's t_eval_modes entry
mkdir -p gpurun_out
for lib in "$@"; do
  B200DAE_LIB=$PWD/$lib timeout 300 python bench.py --t-eval-points 1000 --no-configs --no-strong --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); t=d['t_eval_modes'][0]
print(json.dumps({'lib':'$lib','ms':t['ms_per_launch'],'hbm_gbs':t['hbm_gbs'],'kernel_ms_no_eval':d['roofline']['kernel_ms']}))" | tee -a gpurun_out/ab_teval.jsonl
done
