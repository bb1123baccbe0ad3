"""Unified-memory traffic of the Level-1/2 chains under LD_PRELOAD (BASELINE config 3), counted with CUPTI (tests/drivers/um_counters.c):
the CG chain and the Level-1 chain run with a short and a long iteration count; traffic that does not grow with the count is the
one-off first-touch migration.  Run on a GPU box; prints one summary line per chain (tools only)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_preload import build_driver  # noqa: E402
from helpers import find_openblas  # noqa: E402

LIB = os.path.join(ROOT, "libgpublas_b200", "libb200blas.so")
BUILD = os.path.join(ROOT, "tests", "drivers", "_build")
helper = os.path.join(BUILD, "libumcount.so")
os.makedirs(BUILD, exist_ok=True)
subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I/usr/local/cuda/include", "-o", helper, os.path.join(ROOT, "tests", "drivers", "um_counters.c"),
                       "-L/usr/local/cuda/lib64", "-lcupti", "-L/usr/local/cuda/lib64/stubs", "-lcuda", "-Wl,-rpath,/usr/local/cuda/lib64"])
ob = find_openblas()
env = dict(os.environ, LD_PRELOAD=helper + " " + LIB, OPENBLAS_CORETYPE="SkylakeX")
env["LD_LIBRARY_PATH"] = os.path.dirname(ob) + ":/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")


def counters(exe, args):
    out = subprocess.run([exe] + [str(a) for a in args], env=env, capture_output=True, text=True, timeout=600)
    um = [l for l in out.stdout.splitlines() if l.startswith("UMCOUNT")]
    res = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
    if not um:
        print("no UMCOUNT line:", out.stdout[-500:], out.stderr[-800:])
        return None, res
    return dict(kv.split("=") for kv in um[0].split()[1:]), res


for name, short, long_ in (("cg_chain", [8192, 5], [8192, 105]), ("l1_chain", [1 << 24, 5], [1 << 24, 105])):
    exe = build_driver(name)
    a, ra = counters(exe, short)
    b, rb = counters(exe, long_)
    if a is None or b is None:
        continue
    grow = {k: int(b[k]) - int(a[k]) for k in a}
    print("%s %s iterations: %s" % (name, short[1], a))
    print("%s %s iterations: %s" % (name, long_[1], b))
    print("%s: 100 more iterations added htod_bytes=%d dtoh_bytes=%d gpu_fault_groups=%d cpu_faults=%d  (%s)" % (
        name, grow["htod_bytes"], grow["dtoh_bytes"], grow["gpu_fault_groups"], grow["cpu_faults"], rb[0] if rb else ""), flush=True)
