# usage: bash scripts/runs/_run_mg.sh N   -- weak + strong scaling lines at N GPUs (and the NCCL sharded_loss test at N = 2)
N=$1
mkdir -p gpurun_out/r2
nvidia-smi topo -m > gpurun_out/r2/topo_$N.txt 2>&1
if [ "$N" = "1" ]; then
  python bench.py --scaling strong --steps 200 --warmup 20 > gpurun_out/r2/strong_1.json 2> gpurun_out/r2/strong_1.err
  tail -c 400 gpurun_out/r2/strong_1.json
  exit 0
fi
if [ "$N" = "2" ]; then
  python -m pytest tests/test_dist_cuda.py -m gpu -q > gpurun_out/r2/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_nccl.log
  tail -6 gpurun_out/r2/pytest_nccl.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/r2/weak_$N.json 2> gpurun_out/r2/weak_$N.err
tail -c 300 gpurun_out/r2/weak_$N.json; tail -2 gpurun_out/r2/weak_$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --scaling strong --steps 200 --warmup 20 > gpurun_out/r2/strong_$N.json 2> gpurun_out/r2/strong_$N.err
tail -c 300 gpurun_out/r2/strong_$N.json; tail -2 gpurun_out/r2/strong_$N.err
