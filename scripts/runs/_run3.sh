set -x
mkdir -p gpurun_out/r2
python tests/tools/gpu_debug_golden.py b2_24x32 > gpurun_out/r2/dbg_golden.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest2.log
grep -E "passed|failed|FAILED|rc=" gpurun_out/r2/pytest2.log | tail -30
cat gpurun_out/r2/dbg_golden.log
KERNEL=k_photo_bwd bash scripts/run_variants.sh > gpurun_out/r2/variants_bwd1.log 2>&1
cat gpurun_out/r2/variants_bwd1.log
