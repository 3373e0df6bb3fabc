#!/bin/bash
# A/B of every variant library under scipy_dae_b200/variants/ (scripts/build_variant.py); one JSON line each
mkdir -p gpurun_out
out=gpurun_out/ab_$(date +%H%M%S).jsonl
for pass in 1 2; do
  for lib in scipy_dae_b200/variants/libb200dae_*.so; do
    B200DAE_LIB=$PWD/$lib timeout 120 python scripts/ab_kernel.py "$@" 2>&1 | tail -1 | tee -a $out
  done
done
if [ -n "$AB_FASTMATH_LIB" ]; then
  B200DAE_LIB=$PWD/$AB_FASTMATH_LIB timeout 300 python -m pytest tests/test_gpu_fastmath.py -q -m gpu 2>&1 | tail -15
fi
