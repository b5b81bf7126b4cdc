mkdir -p gpurun_out/r2
for v in base pf4 pf2 pf1 pf4u2; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --profile --kernel 3 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var26.log 2>&1
cat gpurun_out/r2/var26.log
