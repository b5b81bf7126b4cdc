mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r2/pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest13.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2/pytest13.log | tail -6
for v in base nokin; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --profile --kernel 3 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var29.log 2>&1
cat gpurun_out/r2/var29.log
