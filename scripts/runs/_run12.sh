mkdir -p gpurun_out/r2
for v in base S00 S01; do echo $v; COLVO_LIB=$PWD/build/variants/lib_$v.so python bench.py --profile --steps 100 --warmup 10 2>/dev/null | tail -1; done > gpurun_out/r2/variants_S.log
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest5.log
tail -3 gpurun_out/r2/pytest5.log
python bench.py --steps 300 --warmup 20 > gpurun_out/r2/bench2.json 2> gpurun_out/r2/bench2.err; tail -1 gpurun_out/r2/bench2.json | cut -c1-1500
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2/launches2.csv python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_photo_fwd|k_warp_stats|k_photo_bwd" -s 3 -c 3 -f -o gpurun_out/r2/prof_all_v2 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/r2/*.ncu-rep
