mkdir -p gpurun_out/r2
python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_photo_bwd" -s 1 -c 1 -f -o gpurun_out/r2/prof_bwd_v3 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/r2/*.ncu-rep
