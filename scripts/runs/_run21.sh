mkdir -p gpurun_out/r2
for v in tma_h4m5 h4m6 h4m4 h6m3 h2m10 h4m5gp h4m4pk; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel 2 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var21.log 2>&1
cat gpurun_out/r2/var21.log
