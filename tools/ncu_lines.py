#!/usr/bin/env python
"""Stall samples of an `ncu --set full --import-source on` capture aggregated per CUDA source line.
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o built with -lineinfo> [top N] [kernel-name substring]
The SASS page of the report gives samples per instruction address; nvdisasm -g of the cubin gives the line of each address."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
kfilter = sys.argv[4] if len(sys.argv) > 4 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# per function: ordered (line, file) per instruction
funcs, cur, line = {}, None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", ln)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        funcs[cur].append(line)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kname = rows[0][1]
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
# pick the function with the same instruction count
cands = [f for f, l in funcs.items() if len(l) == len(body) and kfilter in f]
if not cands:
    print("no function with", len(body), "instructions; have", {f: len(l) for f, l in funcs.items()}); sys.exit(1)
lines = funcs[cands[0]]
agg = collections.defaultdict(lambda: [0, 0, 0])
for (ln, r) in zip(lines, body):
    a = agg[ln]
    a[0] += int(float(r[ci["# Samples"]] or 0))
    a[1] += int(float(r[ci["Instructions Executed"]] or 0))
    a[2] += int(float(r[ci["L1 Wavefronts Shared"]] or 0)) if "L1 Wavefronts Shared" in ci else 0
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values()); totw = sum(a[2] for a in agg.values()) or 1
print(kname[:100]); print("samples", tot, "warp instructions", toti, "shared wavefronts", totw)
srcs = {}
def text(f, n):
    for root in ("transformerupscaler_b200/csrc", "transformerupscaler_b200/csrc/tc"):
        p = os.path.join(root, f)
        if os.path.exists(p):
            if p not in srcs: srcs[p] = open(p).read().splitlines()
            return srcs[p][n - 1].strip()[:110] if n - 1 < len(srcs[p]) else ""
    return ""
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-28s %5.1f%% samples %5.1f%% instr %5.1f%% smem-wf | %s" % ("%s:%d" % ln if ln else "?", 100 * a[0] / tot, 100 * a[1] / toti, 100 * a[2] / totw, text(*ln) if ln else ""))
