#!/bin/bash
# time every tuning variant on the B200 (bench.py --profile: device-timed fwd+bwd, no e2e / CPU legs);
# KERNEL=<regex> additionally prints that kernel's isolated duration (ncu, serialised)
for f in build/variants/lib_*.so; do
  echo "== $f"
  COLVO_LIB=$PWD/$f python bench.py --profile --steps 40 --warmup 10 2>&1 | tail -1
  if [ -n "$KERNEL" ]; then
    COLVO_LIB=$PWD/$f ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$KERNEL" -c 3 --csv --log-file /tmp/l.csv python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
    python scripts/launch_summary.py /tmp/l.csv | tail -2 | head -1
  fi
done
