mkdir -p gpurun_out/r2
for v in S00 S01; do
COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --steps 100 --warmup 10 2>/dev/null | tail -1
done
python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1 && \
COLVO_LIB=$PWD/build/variants/lib_S00.so ncu --set full --clock-control none --import-source on -k regex:k_photo_bwd -s 1 -c 1 -f -o gpurun_out/r2/prof_bwd_S00 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
COLVO_LIB=$PWD/build/variants/lib_S01.so ncu --set full --clock-control none --import-source on -k regex:k_photo_bwd -s 1 -c 1 -f -o gpurun_out/r2/prof_bwd_S01 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_photo_fwd|k_warp_stats|k_photo_bwd" -s 3 -c 3 -f -o gpurun_out/r2/prof_all_v1 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/r2/*.ncu-rep
