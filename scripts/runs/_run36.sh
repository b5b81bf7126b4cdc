mkdir -p gpurun_out/r2
for rep in 1 2; do for v in base ppt2 ppt3 ppt4 ppt5 ppt6; do echo -n "$v: "; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 200 python bench.py --profile --steps 400 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4))"; done; done | tee gpurun_out/r2/var36b.log
for v in base ppt4; do echo -n "sweep $v: "; COLVO_LIB=$PWD/build/variants/lib_$v.so timeout 200 python bench.py --config 5 --profile --steps 10 --warmup 3 2>/dev/null | tail -1; done | tee -a gpurun_out/r2/var36b.log
