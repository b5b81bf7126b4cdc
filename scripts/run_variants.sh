#!/bin/bash
# time every tuning variant on the B200 (bench.py --profile: device-timed fwd+bwd, no e2e / CPU legs)
for f in build/variants/lib_*.so; do
  echo "== $f"
  COLVO_LIB=$PWD/$f python bench.py --profile --steps 40 --warmup 10 2>&1 | tail -1
done
