set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest1.log
tail -15 gpurun_out/r2/pytest1.log
python bench.py --profile --steps 200 --warmup 20 2> gpurun_out/r2/bench1.err | tail -1
python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2/launches1.csv python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
python scripts/launch_summary.py gpurun_out/r2/launches1.csv | tail -15
