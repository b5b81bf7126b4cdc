set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest0.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest0.log
python bench.py > gpurun_out/r2/bench0.json 2> gpurun_out/r2/bench0.err
python scripts/exp_subbatch.py > gpurun_out/r2/subbatch.jsonl 2> gpurun_out/r2/subbatch.err
KERNEL=k_photo_bwd bash scripts/run_variants.sh > gpurun_out/r2/ablation_bwd.log 2>&1
nvidia-smi topo -m > gpurun_out/r2/topo.txt 2>&1
lscpu > gpurun_out/r2/lscpu.txt 2>&1
tail -3 gpurun_out/r2/pytest0.log; cat gpurun_out/r2/bench0.json | cut -c1-600; cat gpurun_out/r2/subbatch.jsonl; cat gpurun_out/r2/ablation_bwd.log
