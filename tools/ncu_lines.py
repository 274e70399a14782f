"""Per-source-line share of executed warp instructions and stall samples for one kernel of an .ncu-rep.

usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex lib.so [min_pct]
Joins `ncu --page source --print-source sass` (per-instruction counters, in address order) with `nvdisasm -g`
(line info) by instruction index inside the kernel.
"""
import csv, io, re, subprocess, sys, tempfile, os, glob

rep, kre, lib = sys.argv[1:4]
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.5
mangled_has = sys.argv[5].split(",") if len(sys.argv) > 5 else []   # extra substrings the mangled name must contain (template args)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# may contain several launches: keep the first
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
start = [i for i in start if all(x in rows[i][1] for x in (sys.argv[5].split(",") if len(sys.argv) > 5 else []))] or start
nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > start[0]]
seg = rows[start[0]:(nxt[0] if nxt else len(rows))]
kname = seg[0][1]
hdr = seg[1]
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
inst = [(r[1].strip(), int(r[ci] or 0), int(r[cs] or 0)) for r in seg[2:] if len(r) > ci]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
short = kname.split("(")[0].split("<")[0].split("::")[-1].split()[-1]
lines = None
for f in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-g", "-c", f], capture_output=True, text=True).stdout
    if short not in txt:
        continue
    cur, out, active = None, [], False
    for l in txt.splitlines():
        if l.startswith(".text."):
            active = short in l and all(x in l for x in mangled_has) and (not out)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            out.append(cur)
    if out:
        lines = out
        break
assert lines, "kernel not found in cubins"
n = min(len(lines), len(inst))
if len(lines) != len(inst):
    print("# warning: %d sass rows vs %d disassembled" % (len(inst), len(lines)))
agg = {}
tot_i = sum(i for _, i, _ in inst) or 1
tot_s = sum(s for _, _, s in inst) or 1
for k in range(n):
    a = agg.setdefault(lines[k], [0, 0])
    a[0] += inst[k][1]; a[1] += inst[k][2]
src = {}
print("# %s: %d warp-inst, %d samples" % (kname[:80], tot_i, tot_s))
for (key, (i, s)) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if 100.0 * i / tot_i >= minpct or 100.0 * s / tot_s >= minpct:
        f, ln = key or ("?", 0)
        if f not in src:
            p = os.path.join(os.path.dirname(__file__), "..", "elprep_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
        print("%5.1f%% inst %5.1f%% stall  %s:%d  %s" % (100.0 * i / tot_i, 100.0 * s / tot_s, f, ln, text))
