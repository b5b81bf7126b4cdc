mkdir -p gpurun_out/r2
python tests/tools/gpu_tma_dbg.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r2/pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest11.log
grep -E "passed|failed|FAILED|rc=|Error" gpurun_out/r2/pytest11.log | tail -8
for v in base fwd_notma; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --profile --kernel 1 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var23.log 2>&1
cat gpurun_out/r2/var23.log
