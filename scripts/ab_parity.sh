#!/bin/bash
# Robertson parity tests against every variant library (scripts/build_variant.py ... --stages)
for lib in scipy_dae_b200/variants/libb200dae_*.so; do
  echo "=== $lib"
  B200DAE_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "robertson" 2>&1 | grep -E "^E  +AssertionError|passed|failed" | cut -c1-300
done
