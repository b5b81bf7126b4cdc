mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x > gpurun_out/r2/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest6.log
tail -3 gpurun_out/r2/pytest6.log
for v in base nostream; do echo "== $v"; for w in 1 2 3; do COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel $w --steps 100 --warmup 10 2>/dev/null | tail -1; done; done > gpurun_out/r2/var14.log 2>&1
cat gpurun_out/r2/var14.log
