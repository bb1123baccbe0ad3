"""Device-resident timing of the banded / packed / Hermitian Level-2 kernels (level2_struct.cu) with CUDA events.
Algorithmic bytes: the stored part of the matrix once (+ vectors); rank updates read and write it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g

lib = g.load(); g.use_torch_stream(); g.set_sync(False)
import json
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6553.9     # the round's first measurement, used when the driver's file is absent


def time_call(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
n = 32768
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
ap = torch.rand(n * (n + 1) // 2, dtype=torch.float64, device="cuda") * 1e-5
pk = 8.0 * n * (n + 1) / 2
rows.append(("dspmv U n=32768", pk, time_call(lambda: g.call("dspmv_", "U", n, 1.0, ap, x, 1, 0.0, y, 1))))
rows.append(("dspmv L n=32768", pk, time_call(lambda: g.call("dspmv_", "L", n, 1.0, ap, x, 1, 0.0, y, 1))))
rows.append(("dtpmv UNN n=32768", pk, time_call(lambda: g.call("dtpmv_", "U", "N", "N", n, ap, x, 1))))
rows.append(("dtpmv UTN n=32768", pk, time_call(lambda: g.call("dtpmv_", "U", "T", "N", n, ap, x, 1))))
rows.append(("dspr U n=32768", 2 * pk, time_call(lambda: g.call("dspr_", "U", n, 1e-9, y, 1, ap))))
rows.append(("dspr2 L n=32768", 2 * pk, time_call(lambda: g.call("dspr2_", "L", n, 1e-9, y, 1, x, 1, ap))))
x.fill_(1.0)
# diagonal of the packed upper triangle at j(j+1)/2 + j: make the solve well conditioned
idx = torch.arange(n, device="cuda", dtype=torch.int64); ap[idx * (idx + 1) // 2 + idx] = 2.0
rows.append(("dtpsv UNN n=32768", pk, time_call(lambda: g.call("dtpsv_", "U", "N", "N", n, ap, x, 1), reps=3, warm=1)))
x.fill_(1.0)
rows.append(("dtpsv UTN n=32768", pk, time_call(lambda: g.call("dtpsv_", "U", "T", "N", n, ap, x, 1), reps=3, warm=1)))
del ap
nb, kl, ku = 1 << 22, 63, 64
ab = torch.rand((nb, kl + ku + 1), dtype=torch.float64, device="cuda")     # memory = column-major (kl+ku+1) x nb
xb = torch.rand(nb, dtype=torch.float64, device="cuda"); yb = torch.zeros(nb, dtype=torch.float64, device="cuda")
bb = 8.0 * nb * (kl + ku + 1)
rows.append(("dgbmv N n=2^22 kl=63 ku=64", bb, time_call(lambda: g.call("dgbmv_", "N", nb, nb, kl, ku, 1.0, ab, kl + ku + 1, xb, 1, 0.0, yb, 1))))
rows.append(("dgbmv T n=2^22 kl=63 ku=64", bb, time_call(lambda: g.call("dgbmv_", "T", nb, nb, kl, ku, 1.0, ab, kl + ku + 1, xb, 1, 0.0, yb, 1))))
k = 127
rows.append(("dsbmv U n=2^22 k=127", bb, time_call(lambda: g.call("dsbmv_", "U", nb, k, 1.0, ab, k + 1, xb, 1, 0.0, yb, 1))))
ab[:, k] = 4.0 * k
xb.fill_(1.0)
rows.append(("dtbsv UNN n=2^18 k=127", 8.0 * (1 << 18) * 128, time_call(lambda: g.call("dtbsv_", "U", "N", "N", 1 << 18, k, ab, k + 1, xb, 1), reps=3, warm=1)))
del ab
m = 16384
Z = torch.rand((m, m), dtype=torch.complex128, device="cuda"); zx = torch.rand(m, dtype=torch.complex128, device="cuda"); zy = torch.rand(m, dtype=torch.complex128, device="cuda")
rows.append(("zhemv U n=16384", 16.0 * m * (m + 1) / 2, time_call(lambda: g.call("zhemv_", "U", m, 1.0 + 0j, Z, m, zx, 1, 0j, zy, 1))))
rows.append(("zher2 L n=16384", 2 * 16.0 * m * (m + 1) / 2, time_call(lambda: g.call("zher2_", "L", m, 1e-9 + 0j, zx, 1, zy, 1, Z, m))))
rows.append(("zgeru n=16384", 2 * 16.0 * m * m, time_call(lambda: g.call("zgeru_", m, m, 1e-9 + 0j, zx, 1, zy, 1, Z, m))))
del Z
nn = 32768
Sf = torch.rand((nn, nn), dtype=torch.float64, device="cuda"); sx = torch.rand(nn, dtype=torch.float64, device="cuda"); sy = torch.zeros(nn, dtype=torch.float64, device="cuda")
for ul in "UL":
    rows.append(("dsymv %s n=32768" % ul, 8.0 * nn * (nn + 1) / 2, time_call(lambda: g.call("dsymv_", ul, nn, 1.0, Sf, nn, sx, 1, 0.0, sy, 1))))
del Sf
ab = torch.rand((nb, k + 1), dtype=torch.float64, device="cuda")
rows.append(("dsbmv L n=2^22 k=127", 8.0 * nb * (k + 1), time_call(lambda: g.call("dsbmv_", "L", nb, k, 1.0, ab, k + 1, xb, 1, 0.0, yb, 1))))
for name, byts, ms in rows:
    print(f"{name:30s} {ms:9.4f} ms  {byts/ms/1e6:8.1f} GB/s  {byts/ms/1e6/PEAK*100:5.1f}% of measured HBM peak", flush=True)
