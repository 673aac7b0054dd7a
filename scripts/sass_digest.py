"""SASS digest of lib/libsfb200.so: per kernel, how often the instructions that matter appear (no GPU needed).
    python scripts/sass_digest.py > profiles/r02_sass_digest.txt
DMMA = FP64 tensor core (mma.sync.m8n8k4.f64), UBLKCP = cp.async.bulk (TMA engine), UTMALDG = cp.async.bulk.tensor,
SYNCS = mbarrier, REDG/RED = red.global.add (fire-and-forget atomic adds, also over NVLink into peer images),
LDGSTS = cp.async, DFMA / DADD / DMUL = FP64 vector pipe."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sparse-matrix-factorization-library_b200", "lib", "libsfb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
WATCH = ["DMMA", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "REDG", "RED.", "ATOMG", "LDGSTS", "DFMA", "DADD", "DMUL", "MUFU", "LDS", "STS", "LDG", "STG",
         "UTCHMMA", "LDTM", "WARPSYNC", "BAR.SYNC", "NANOSLEEP", "MEMBAR", "ERRBAR", "FENCE"]
arch = re.findall(r"arch = (sm_\w+)", sass)
print("library:", os.path.relpath(lib, ROOT))
print("architectures in the fatbin:", sorted(set(arch)))
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"sfb200::", "", cur)
        cur = cur.split("(")[0]
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
regs = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name); name = re.sub(r"sfb200::", "", name).split("(")[0]
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and name:
        regs[name] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    r = regs.get(k, ("?", "?", "?"))
    print("\n%s\n    instructions %d, registers %s, static smem %s B, local (spill) %s B" % (k, c["total"], r[0], r[1], r[2]))
    print("    " + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]))
print("\nwhole library: " + "  ".join("%s %d" % (w, tot[w]) for w in WATCH if tot[w]))
ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
print("\nlinked CUDA libraries:", sorted(set(re.findall(r"(lib(?:cu|nccl|nv)\w*)\.so", ldd))))
