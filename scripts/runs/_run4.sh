set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest3.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest3.log | tail -30
grep -E "^E  .*(Error|assert)" gpurun_out/r2/pytest3.log | cut -c1-200 | head -20
./build/ffma2_bench > gpurun_out/r2/ffma2_bench.txt 2>&1
cat gpurun_out/r2/ffma2_bench.txt
python bench.py --profile --steps 200 --warmup 20 2>/dev/null | tail -1
