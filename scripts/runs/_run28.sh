mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest12.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest12.log | tail -8
python bench.py --config 4 --steps 30 --warmup 5 > gpurun_out/r2/bench5_c4.json 2> gpurun_out/r2/bench5_c4.err; tail -c 300 gpurun_out/r2/bench5_c4.json; tail -3 gpurun_out/r2/bench5_c4.err
python bench.py --steps 300 --warmup 20 > gpurun_out/r2/bench5_c2.json 2> gpurun_out/r2/bench5_c2.err; tail -c 300 gpurun_out/r2/bench5_c2.json; tail -3 gpurun_out/r2/bench5_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2/launches_c4.csv python bench.py --config 4 --profile --steps 2 --warmup 1 > /dev/null 2>&1
