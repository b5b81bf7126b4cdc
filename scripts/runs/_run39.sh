python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "warp_aggregated or parity_small or host_stepper or graphed" 2>&1 | grep -E "passed|failed|FAILED|merged" | tail -6
python bench.py --steps 300 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['config']['eager_ms_per_step'],4), d['config']['without_image_gradient']['ms_per_step'], d['config']['warp_aggregated_scatter']['ms_per_step'], d['roofline']['traffic_stale'])"
