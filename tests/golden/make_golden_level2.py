"""Generates tests/golden/level2_struct_openblas.npz: outputs of the CPU BLAS of this image (OpenBLAS 0.3.15, Fortran symbols)
for a slice of the shared Level-2 case list (tests/l2x.py: every 7th case of sizes (5, 33), precisions d and z).  The
reference ships no Level-1/2 tester and no expected values, so this pins the oracle and the GPU tests to
what the library behind the reference's interposer computed on this input, without needing that library at test time.
Keys: "<p>/<index>" -> the output array of case <index> of l2x.cases(p, sizes=(5, 33)); "<p>/tags" -> the case tags (so a change
of the case list is noticed).  Run from the repository root:  python tests/golden/make_golden_level2.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import l2x  # noqa: E402
from helpers import f77, load_openblas  # noqa: E402

ob = load_openblas()
out = {}
for p in "dz":
    cs = l2x.cases(p, sizes=(5, 33))[::7]
    out[p + "/tags"] = np.array([c.tag for c in cs])
    for i, c in enumerate(cs):
        args = c.fresh_args()
        f77(ob, c.name + "_", *args)
        out["%s/%d" % (p, i)] = args[c.out]
np.savez_compressed(os.path.join(HERE, "level2_struct_openblas.npz"), **out)
print("cases:", {p: len(out[p + "/tags"]) for p in "dz"})
