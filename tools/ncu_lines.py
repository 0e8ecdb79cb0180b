#!/usr/bin/env python
"""Development aid: attribute the warp-stall samples of an ncu report to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [top]

ncu's source page gives samples per SASS instruction (in program order); nvdisasm -g gives the source line of every SASS
instruction of the same kernel (the object must be the one the profiled library was linked from, compiled with -lineinfo).
Prints the heaviest (file:line, innermost inlining level) with their share of samples and of executed instructions.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, pick = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
samples = [int(r[ix["# Samples"]] or 0) for r in data]
execd = [int(r[ix["Instructions Executed"]] or 0) for r in data]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
    cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith("//--------------------- .text."):
        inside = pick in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
n = min(len(lines), len(data))
if len(lines) != len(data):
    print(f"warning: {len(lines)} disassembled instructions vs {len(data)} in the report", file=sys.stderr)
agg_s, agg_e = collections.Counter(), collections.Counter()
for i in range(n):
    agg_s[lines[i]] += samples[i]
    agg_e[lines[i]] += execd[i]
ts, te = sum(samples), sum(execd)
print(f"{'file:line':40s} samples%  executed%")
for key, s in agg_s.most_common(top):
    print(f"{key[0] + ':' + str(key[1]):40s} {100 * s / ts:7.2f}  {100 * agg_e[key] / te:8.2f}")
byfile = collections.Counter()
for k, s in agg_s.items():
    byfile[k[0]] += s
print({k: round(100 * v / ts, 1) for k, v in byfile.most_common()})
