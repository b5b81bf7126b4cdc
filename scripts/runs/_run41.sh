mkdir -p gpurun_out/r2
python scripts/e2e_chunks.py > gpurun_out/r2/e2e_chunks.log 2>&1; cat gpurun_out/r2/e2e_chunks.log
