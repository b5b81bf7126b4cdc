mkdir -p gpurun_out/r2
for v in base sw96 sw48 sw32 sw16; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --config 5 --profile --steps 10 --warmup 3 2>/dev/null | tail -1; done > gpurun_out/r2/var27.log 2>&1
cat gpurun_out/r2/var27.log
