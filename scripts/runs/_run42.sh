python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x 2>&1 | grep -E "passed|failed|FAILED" | tail -3
for rep in 1 2; do for v in cur prev; do echo -n "$v: "; L=$PWD/build/variants/lib_prev.so; [ $v = cur ] && L=$PWD/coivo_b200/libcolvo_b200.so; COLVO_LIB=$L python bench.py --profile --steps 400 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))"; done; done
for v in cur prev; do echo -n "c4 $v: "; L=$PWD/build/variants/lib_prev.so; [ $v = cur ] && L=$PWD/coivo_b200/libcolvo_b200.so; COLVO_LIB=$L python bench.py --config 4 --profile --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))"; done
