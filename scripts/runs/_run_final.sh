# final record of a build: tests, smoke, bench lines for the three configs, launch list and full ncu capture of one step per
# config; the captures are digested ON the box (scripts/make_traffic.py, ncu_summary.py, ncu_hot_sass.py) because gpurun_out/ is
# capped at 64 MiB -- only the config-2 report itself travels back
O=gpurun_out/final
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
grep -E "passed|failed|FAILED|rc=" $O/pytest.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
python bench.py --steps 500 --warmup 20 > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 200 $O/bench_c2.json
python bench.py --config 4 --steps 40 --warmup 5 > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 200 $O/bench_c4.json
python bench.py --config 5 --steps 10 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 200 $O/bench_c5.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 200 $O/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_c2.csv python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
T=/tmp/colvo_ncu; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 9 -c 9 -f -o $T/step_c2 python bench.py --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 9 -c 9 -f -o $T/step_c4 python bench.py --config 4 --profile --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_" -s 34 -c 34 -f -o $T/step_c5 python bench.py --config 5 --profile --steps 2 --warmup 1 > /dev/null 2>&1
cp profiles/traffic.json $O/traffic.json 2>/dev/null
for c in 2 4 5; do python scripts/make_traffic.py --config $c --out $O/traffic.json $T/step_c$c.ncu-rep > $O/traffic_c$c.txt 2>&1; done
python scripts/ncu_summary.py $T/step_c2.ncu-rep > $O/ncu_full_summary_c2.txt 2>&1
python scripts/ncu_summary.py $T/step_c4.ncu-rep > $O/ncu_full_summary_c4.txt 2>&1
for k in k_warp_stats k_photo_fwd k_photo_bwd; do python scripts/ncu_hot_sass.py $T/step_c2.ncu-rep $k 24; done > $O/ncu_hot_sass_c2.txt 2>&1
cp $T/step_c2.ncu-rep $O/
ls -la $O; du -sh gpurun_out
