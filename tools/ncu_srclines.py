"""Per-source-line share of executed warp instructions of one kernel in an .ncu-rep captured with --import-source on.
usage: python tools/ncu_srclines.py rep.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys

rep, kre = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
sec, hdr, out, fn = None, None, [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": sec = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name": fn = r[1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = dict(zip(hdr, r))
        try: ie = int(d["Instructions Executed"])
        except ValueError: continue
        out.append((ie, sec, d["Line No"], " ".join(r[1].split())[:140]))
tot = sum(o[0] for o in out)
print("kernel:", fn); print("warp instructions attributed to source lines:", tot)
out.sort(reverse=True); acc = 0
for ie, s, l, t in out[:top]:
    acc += ie
    print(f"{ie / tot * 100:5.1f}% {acc / tot * 100:5.1f}%  {s}:{l}  {t}")
