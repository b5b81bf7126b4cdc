mkdir -p gpurun_out/r2
for v in base neartaps nored neartaps_nored; do echo "== $v"; for w in 1 2 3; do COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --kernel $w --steps 100 --warmup 10 2>/dev/null | tail -1; done; done > gpurun_out/r2/abl_neartaps.log 2>&1
cat gpurun_out/r2/abl_neartaps.log
