mkdir -p gpurun_out/r2
for rep in 1 2; do for v in base keep zero8 zero2 h2m10; do echo -n "$v: "; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 200 python bench.py --profile --steps 400 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))"; done; done | tee gpurun_out/r2/var36c.log
