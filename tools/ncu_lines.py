"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py rep.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) > 7 and r[2] == "-":
        d = dict(zip(hdr, r))
        recs.append((cur_file, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0), d))
tot = sum(x[3] for x in recs) or 1
toti = sum(x[4] for x in recs) or 1
print(f"total samples {tot}, warp-instructions {toti}")
for f, ln, src, smp, ins, d in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{100*smp/tot:5.1f}% smp {100*ins/toti:5.1f}% ins  {f}:{ln:<4d} {src[:110]}")
