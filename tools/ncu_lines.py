#!/usr/bin/env python
"""Per-source-line share of a kernel's executed warp instructions and stall samples.

    python tools/ncu_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [launch index]

The ncu source page (SASS view, CSV) carries one row per instruction with its counters but no line numbers; nvdisasm -g
of the object's cubin carries `//## File "...", line N` markers.  Both list the function's instructions in address order,
so they are joined by offset."""
import csv, io, os, re, subprocess, sys, tempfile

rep, obj, kern = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, infn = None, False
for l in dis:
    if l.startswith(".text."):
        infn = kern in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True,
                     text=True).stdout
blocks = raw.split('"Kernel Name"')
ncu_name = os.environ.get("NCU_NAME", kern)      # demangled name in the report when it differs from the mangled one
blocks = [b for b in blocks[1:] if ncu_name in b.splitlines()[0]]
rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[which])))
h = next(r for r in rows if "Address" in r and "Source" in r)
ia, ii, ismp = h.index("Address"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
body = [r for r in rows[rows.index(h) + 1:] if len(r) > ii and r[ia].startswith("0x")]
base = int(body[0][ia], 16)
agg = {}
ti = ts = 0
for r in body:
    off = int(r[ia], 16) - base
    key = line_of.get(off)
    n, s = int(r[ii] or 0), int(r[ismp] or 0)
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
    ti += n
    ts += s
src = {}
print(f"kernel {kern}: {ti} warp instructions, {ts} stall samples")
print("| line | instr % | samples % | source |\n|---|---|---|---|")
for key, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    text = ""
    if key:
        f = next((os.path.join(d, key[0]) for d, _, fs in os.walk(os.path.dirname(os.path.abspath(obj)) + "/..") if key[0] in fs), None)
        if f:
            if f not in src:
                src[f] = open(f).read().splitlines()
            text = src[f][key[1] - 1].strip()[:110]
    print(f"| {key[0] + ':' + str(key[1]) if key else '?'} | {100 * n / max(ti, 1):.1f} | {100 * s / max(ts, 1):.1f} | `{text}` |")
