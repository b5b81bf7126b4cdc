mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "parity_small or config2_full or geometric or identity or determinism" > gpurun_out/r2/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest7.log
tail -3 gpurun_out/r2/pytest7.log
for v in base merge_m2 merge_h4_m5 nomerge nomerge_h4_m5; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel 2 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var17.log 2>&1
cat gpurun_out/r2/var17.log
