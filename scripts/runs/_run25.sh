mkdir -p gpurun_out/r2
for v in base fwd_noearly fwd_allearly fwd_notma; do echo "== $v"; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 300 python bench.py --profile --kernel 1 --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/var25.log 2>&1
cat gpurun_out/r2/var25.log
ncu --set full --clock-control none --import-source on -k regex:"k_photo_fwd" -s 1 -c 1 -f -o gpurun_out/r2/prof_fwd_tma python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
