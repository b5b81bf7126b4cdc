COLVO_LIB=$PWD/build/variants/lib_probe.so python tests/tools/gpu_tma_dbg.py 2>&1 | tail -8
