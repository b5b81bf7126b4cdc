python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED" | tail -3
for rep in 1 2; do for v in base prev; do echo -n "$v: "; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --steps 400 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value'],1), round(d['ms_per_step'],4), round(d['config']['eager_ms_per_step'],4), round(d['config']['without_image_gradient']['ms_per_step'],4), round(d['e2e']['value'],1))"; done; done
