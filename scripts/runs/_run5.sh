mkdir -p gpurun_out/r2
python tests/tools/gpu_debug_golden.py b1_16x24_detach 2>&1 | tail -40
python tests/tools/gpu_debug_case.py 3 16 24 1 2 2>&1 | tail -30
