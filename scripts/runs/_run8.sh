COLVO_LIB=$PWD/build/lib_old.so python tests/tools/gpu_dump.py 2 16 24 1 1 /tmp/a.pt > /dev/null
python tests/tools/gpu_dump.py 2 16 24 1 1 /tmp/b.pt > /dev/null
python tests/tools/gpu_dump_cmp.py /tmp/a.pt /tmp/b.pt map
