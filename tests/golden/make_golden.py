"""Generates tests/golden/gemm_c_fixture.json: the known answer of the reference's own micro-benchmark
fixture /root/reference/tests/c/gemm.c:29-35 (A[i + j*m] = i, B[i + j*k] = j, alpha = 1, beta = 0,
m = n = k) restated in f64, where C[i,j] = k*i*j holds EXACTLY (k*i*j <= 1.07e9 < 2^53 at n = 1024).
No library is involved: the expected values come from integer arithmetic."""
import json
import os

import numpy as np

cases = []
for n in (4, 64, 256):
    i = np.arange(n, dtype=np.int64)
    C = (n * np.outer(i, i)).astype(np.float64)
    samples = [[int(r), int(c), float(C[r, c])] for (r, c) in [(0, 0), (1, 1), (n - 1, n - 1), (n // 2, n - 1), (n - 1, 1)]]
    cases.append({"n": n, "samples": samples, "sum": float(C.sum()),
                  "xor_bits": int(np.bitwise_xor.reduce(np.asfortranarray(C).view(np.uint64).ravel()))})
json.dump({"source": "reference tests/c/gemm.c:29-35, f64 restatement; C[i,j] = n*i*j", "cases": cases},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gemm_c_fixture.json"), "w"), indent=1)
