mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest10.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest10.log | tail -8
python bench.py --steps 300 --warmup 20 > gpurun_out/r2/bench4_c2.json 2> gpurun_out/r2/bench4_c2.err; tail -c 300 gpurun_out/r2/bench4_c2.json; tail -3 gpurun_out/r2/bench4_c2.err
for w in 1 3; do python bench.py --profile --kernel $w --steps 100 --warmup 10 2>/dev/null | tail -1; done
python tests/tools/gpu_fuzz.py 3 60 > gpurun_out/r2/fuzz1.log 2>&1; tail -3 gpurun_out/r2/fuzz1.log
