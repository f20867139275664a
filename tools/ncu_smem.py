"""Per CUDA source line: shared-memory wavefronts (total / excessive) from an ncu report.
usage: python tools/ncu_smem.py rep.ncu-rep kernel_regex [top]"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and r[2] == "-":
        d = dict(zip(hdr, r))
        recs.append((cur, int(r[0]), r[1].strip(), int(d["L1 Wavefronts Shared"] or 0),
                     int(d["L1 Wavefronts Shared Excessive"] or 0), int(d["Instructions Executed"] or 0)))
tw = sum(x[3] for x in recs) or 1
te = sum(x[4] for x in recs)
print(f"shared wavefronts {tw}, excessive {te} ({100*te/tw:.1f}%)")
for f, ln, src, w, e, ins in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{100*w/tw:5.1f}% wav  exc {100*e/max(w,1):5.1f}%  {f}:{ln:<4d} {src[:100]}")
