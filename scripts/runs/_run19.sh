mkdir -p gpurun_out/r2
COLVO_LIB=$PWD/build/variants/lib_tma_t0wait.so python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "parity_small or config2_full or identity or determinism or highres_slice" > gpurun_out/r2/pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest9.log
tail -3 gpurun_out/r2/pytest9.log
for v in tma_t0wait tma_allwait notma; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel 2 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var19.log 2>&1
cat gpurun_out/r2/var19.log
