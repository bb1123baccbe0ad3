cd /root/repo
export LD_LIBRARY_PATH=$(dirname $(python -c "import sys; sys.path.insert(0,'tests'); from helpers import find_openblas; print(find_openblas())"))
python -c "
import sys; sys.path.insert(0,'tests')
import test_preload as t
print(t.build_driver('allocs'))"
E=tests/drivers/_build/allocs
echo "--- heuristic=true"
BLAS2CUDA_OPTIONS="heuristic=true" LD_PRELOAD=$PWD/libgpublas_b200/libb200blas.so timeout 60 $E > /tmp/o.txt 2>/tmp/e.txt; echo rc=$?; cat /tmp/o.txt; tail -5 /tmp/e.txt
echo "--- heuristic=true to tty-like (stdbuf -oL)"
BLAS2CUDA_OPTIONS="heuristic=true" LD_PRELOAD=$PWD/libgpublas_b200/libb200blas.so timeout 60 stdbuf -o0 $E; echo rc=$?
