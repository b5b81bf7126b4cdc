mkdir -p gpurun_out/r2
for v in notma tma_h4m5 notma_h4m5 tma_h8m2; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel 2 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var20.log 2>&1
cat gpurun_out/r2/var20.log
