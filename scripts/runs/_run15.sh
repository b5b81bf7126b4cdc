mkdir -p gpurun_out/r2
for v in head geo_smem geo_m2 geo_m2_packed geo_m2_gp; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel 2 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var15.log 2>&1
cat gpurun_out/r2/var15.log
