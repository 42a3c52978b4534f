#!/usr/bin/env python3
"""Maps the per-instruction stall samples of an ncu report (--set full, --import-source on) to CUDA source lines with the line table
of the cubin (nvdisasm -g), because `ncu --page source --csv --print-source cuda` carries no metrics.

  python tools/ncu_hot_lines.py REPORT.ncu-rep OBJECT.o MANGLED_KERNEL_SUBSTRING [top=40]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=tmp, capture_output=True, text=True).stdout.splitlines()
# offset -> (file, line) of the innermost frame and of the outermost (kernel-body) frame
inner, outer = {}, {}
cur_in = cur_out = None
active = False
for ln in dis:
    if ln.startswith("\t.section") or ln.startswith(".section"):
        active = (".text." in ln) and (kern in ln)
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = os.path.basename(m.group(1)), int(m.group(2)), m.group(3)
        if "inlined at" in rest:
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', rest)
            cur_in = (f, l)
            cur_out = (os.path.basename(m2.group(1)), int(m2.group(2)))
        else:
            cur_in = cur_out = (f, l)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_in:
        off = int(m.group(1), 16)
        inner[off] = cur_in
        outer[off] = cur_out
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
base = int(rows[0]["Address"], 16)
by_in, by_out = collections.Counter(), collections.Counter()
inst_in = collections.Counter()
thr_in = collections.Counter()
reasons = collections.defaultdict(collections.Counter)
total = 0
tot_inst = 0
for r in rows:
    off = int(r["Address"], 16) - base
    s = int(r["# Samples"] or 0)
    n = int(r["Instructions Executed"] or 0)
    total += s
    tot_inst += n
    k = inner.get(off, ("?", 0))
    by_in[k] += s
    inst_in[k] += n
    thr_in[k] += int(r["Thread Instructions Executed"] or 0)
    by_out[outer.get(off, ("?", 0))] += s
    for key, v in r.items():
        if key.startswith("stall_") and "Not Issued" not in key and v and int(v):
            reasons[k][key[6:]] += int(v)
print("kernel %s: %d samples, %d warp instructions" % (kern, total, tot_inst))
print("-- innermost source line: samples %, warp-instructions %, lanes per instruction, top stall reasons")
for k, s in by_in.most_common(top):
    rs = ", ".join("%s %d" % kv for kv in reasons[k].most_common(3))
    print("%-18s:%-5d %5.1f%% %5.1f%% %5.1f  %s" % (k[0], k[1], 100.0 * s / total, 100.0 * inst_in[k] / max(tot_inst, 1), thr_in[k] / max(inst_in[k], 1), rs))
print("-- outermost frame (line of the kernel body or of the out-of-line function that the instruction was inlined into): samples %")
for k, s in by_out.most_common(top):
    print("%-18s:%-5d %5.1f%%" % (k[0], k[1], 100.0 * s / total))
