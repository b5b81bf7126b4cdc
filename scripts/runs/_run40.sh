mkdir -p gpurun_out/r2
timeout 1200 python tests/tools/gpu_fuzz.py 11 150 > gpurun_out/r2/fuzz_final.log 2>&1; tail -3 gpurun_out/r2/fuzz_final.log; grep -c merged gpurun_out/r2/fuzz_final.log
