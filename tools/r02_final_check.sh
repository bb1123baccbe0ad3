#!/bin/sh
# final check of the tree on one GPU: build entry smoke, full GPU suite
TAG=${1:-r02final}
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -4 $OUT/${TAG}_gpu_tests.log
