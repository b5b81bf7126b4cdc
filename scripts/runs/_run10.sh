mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest4.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest4.log | tail -30
grep -E "^E  .*(Error|assert)" gpurun_out/r2/pytest4.log | cut -c1-200 | head -20
python bench.py --profile --steps 200 --warmup 20 2>/dev/null | tail -1
