#!/bin/sh
timeout 200 python -m pytest tests/test_zz_level2_struct_gpu.py -m gpu -x -q -p no:cacheprovider -k "struct_fortran or one_pass" 2>&1 | tail -2
