mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "parity_small or config2_full or config4 or flags" > gpurun_out/r2/pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest14.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest14.log | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2/launches3_c2.csv python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2/launches3_c4.csv python bench.py --config 4 --profile --steps 2 --warmup 1 > /dev/null 2>&1
python bench.py --steps 500 --warmup 20 --no-graph --profile | tail -1
python bench.py --steps 300 --warmup 20 > gpurun_out/r2/bench6_c2.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r2/bench6_c2.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['eager_ms_per_step'], d['config']['without_image_gradient']['ms_per_step'])"
