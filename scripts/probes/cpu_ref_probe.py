"""Why did the reference's CPU rate swing 96 / 189 / 275 GFLOP/s between arms and boxes in round 1?  Times the compiled
reference (oracle/_ref) on lap3d 64^3 in fresh processes: with and without torch / CUDA initialised first, and with
8 / 16 / 32 BLAS threads; prints the host state next to each number.  Test infrastructure (uses oracle/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r"""
import sys, os, time
sys.path.insert(0, os.path.join(%r, 'oracle')); sys.path.insert(0, os.path.join(%r, 'sparse-matrix-factorization-library_b200'))
mode, threads = sys.argv[1], int(sys.argv[2])
if mode == 'torch':
    import torch
    torch.zeros(8, device='cuda').sum().item()
import numpy as np
import sfref
from sparseframe_b200 import matrices
n, r, c, x, sym = matrices.generate('lap3d', 64)
ref = sfref.Reference(n, r, c, x, sym, kind='cholesky', blas_threads=threads).analyze()
sup, lsip = ref.array('Super'), ref.array('Lsip')
cc = np.diff(sup).astype(float); m = np.diff(lsip).astype(float) - cc
fl = float((cc ** 3 / 3 + m * cc * cc + m * (m + 1) * cc).sum())
ts = []
for i in range(4):
    ref.factorize(); ts.append(ref.double('factorizeTime'))
print('RESULT', mode, threads, ' '.join('%%.1f' %% (fl / t / 1e9) for t in ts), 'GFLOP/s per step; loadavg', os.getloadavg(), 'affinity', len(os.sched_getaffinity(0)), 'cpus', os.cpu_count())
""" % (ROOT, ROOT)

for mode, threads in (("fresh", 16), ("torch", 16), ("fresh", 8), ("fresh", 32), ("fresh", 16)):
    p = subprocess.run([sys.executable, "-c", CHILD, mode, str(threads)], capture_output=True, text=True)
    for ln in p.stdout.splitlines():
        if ln.startswith("RESULT"):
            print(ln, flush=True)
    if p.returncode != 0:
        print("FAILED", mode, threads, p.stderr[-500:], flush=True)
try:
    print(subprocess.check_output("lscpu | grep -E 'Model name|Socket|Core|Thread|NUMA|MHz'; uptime; nproc", shell=True, text=True))
except Exception as e:
    print(e)
