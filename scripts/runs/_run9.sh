COLVO_LIB=$PWD/build/variants/lib_dbg.so python tests/tools/gpu_dbg_bwd.py 2 16 24 1 1
