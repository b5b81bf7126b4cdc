for kw in "alpha=1.0" "lcc=False" "smooth_weight=0.0" "lcc_detach=True" "alpha=1.0 lcc_detach=True"; do
echo "=== 3 16 24 1 2 $kw"
COLVO_LIB=$PWD/build/lib_old.so python tests/tools/gpu_dump.py 3 16 24 1 2 /tmp/a.pt $kw > /dev/null
python tests/tools/gpu_dump.py 3 16 24 1 2 /tmp/b.pt $kw > /dev/null
python tests/tools/gpu_dump_cmp.py /tmp/a.pt /tmp/b.pt | grep -E "^gd|^gT"
done
for shape in "1 16 24 1 2" "2 16 24 1 2" "3 16 24 1 1" "3 16 24 1 3" "3 32 24 1 2" "3 16 40 1 2"; do
echo "=== $shape"
COLVO_LIB=$PWD/build/lib_old.so python tests/tools/gpu_dump.py $shape /tmp/a.pt > /dev/null
python tests/tools/gpu_dump.py $shape /tmp/b.pt > /dev/null
python tests/tools/gpu_dump_cmp.py /tmp/a.pt /tmp/b.pt | grep -E "^gd|^gT"
done
