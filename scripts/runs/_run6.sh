mkdir -p gpurun_out/r2
COLVO_LIB=$PWD/build/lib_old.so python tests/tools/gpu_dump.py 3 16 24 1 2 /tmp/a.pt
python tests/tools/gpu_dump.py 3 16 24 1 2 /tmp/b.pt
python tests/tools/gpu_dump_cmp.py /tmp/a.pt /tmp/b.pt
COLVO_LIB=$PWD/build/lib_old.so python tests/tools/gpu_dump.py 1 16 24 2 2 /tmp/a2.pt lcc_detach=True smooth_weight=0.05
python tests/tools/gpu_dump.py 1 16 24 2 2 /tmp/b2.pt lcc_detach=True smooth_weight=0.05
python tests/tools/gpu_dump_cmp.py /tmp/a2.pt /tmp/b2.pt
