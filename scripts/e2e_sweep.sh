#!/bin/bash
# shapes of the host pipeline (bench.py e2e leg); one line each: kernel ms, e2e ms, per-step e2e ms
run() { python bench.py --no-cpu-baseline --t-eval-points 0 "$@" 2>gpurun_out/e2e_sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$*', round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['e2e']['ms_each'])
"; }
run --e2e-chunks 8
run --e2e-chunks 8 --e2e-no-taper
run --e2e-chunks 16
run --e2e-chunks 8 --e2e-streams 3
run --e2e-chunks 16 --e2e-streams 3
run --e2e-chunks 12
