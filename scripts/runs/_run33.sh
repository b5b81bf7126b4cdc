mkdir -p gpurun_out/r2
COLVO_DEBUG=1 bash scripts/runs/_run32.sh 2>&1 | tail -6
COLVO_LIB=$PWD/build/variants/lib_bounds.so timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_frontend.py -m gpu -q > gpurun_out/r2/pytest_bounds.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_bounds.log
grep -E "passed|failed|FAILED|rc=|bounds check" gpurun_out/r2/pytest_bounds.log | tail -8
COLVO_LIB=$PWD/build/variants/lib_bounds.so timeout 900 python tests/tools/gpu_fuzz.py 5 60 > gpurun_out/r2/fuzz_bounds.log 2>&1; tail -2 gpurun_out/r2/fuzz_bounds.log
timeout 900 python tests/tools/gpu_fuzz.py 6 60 > gpurun_out/r2/fuzz2.log 2>&1; tail -2 gpurun_out/r2/fuzz2.log
