mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest8.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest8.log | tail -8
python bench.py --steps 300 --warmup 20 > gpurun_out/r2/bench3_c2.json 2> gpurun_out/r2/bench3_c2.err; tail -c 600 gpurun_out/r2/bench3_c2.json; tail -3 gpurun_out/r2/bench3_c2.err
python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r2/bench3_c4.json 2> gpurun_out/r2/bench3_c4.err; tail -c 600 gpurun_out/r2/bench3_c4.json; tail -3 gpurun_out/r2/bench3_c4.err
python bench.py --config 5 --steps 10 --warmup 3 > gpurun_out/r2/bench3_c5.json 2> gpurun_out/r2/bench3_c5.err; tail -c 600 gpurun_out/r2/bench3_c5.json; tail -3 gpurun_out/r2/bench3_c5.err
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 9 -c 9 -f -o gpurun_out/r2/step_c2 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 9 -c 9 -f -o gpurun_out/r2/step_c4 python bench.py --config 4 --profile --steps 2 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/r2/step_*.ncu-rep
