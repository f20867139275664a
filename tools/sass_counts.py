"""Per-kernel SASS instruction counts of the shipped library (CPU box: cuobjdump only): which kernels use the FP64
tensor instruction (DMMA), TMA bulk copies (UBLKCP / UTMALDG / UTMASTG), cp.async (LDGSTS), mbarriers (SYNCS).
usage: python tools/sass_counts.py [lib.so] > profiles/sass_r2.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "hcore_b200/lib/libhcore_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "UTMALDG", "UTMASTG", "UTCMMA", "LDGSTS", "SYNCS", "SHFL", "LDS", "STS", "BAR", "MUFU"]
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("hcore_b200::", "").replace("void ", "")
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
print(f"SASS instruction counts per kernel, {lib} (cuobjdump -sass, sm_100a)")
print(f"{'kernel':58s} {'instr':>7s} " + " ".join(f"{w:>7s}" for w in WANT))
tot = collections.Counter()
for kname, c in counts.items():
    print(f"{kname[:58]:58s} {total[kname]:7d} " + " ".join(f"{c[w]:7d}" for w in WANT))
    tot.update(c)
print(f"{'TOTAL':58s} {sum(total.values()):7d} " + " ".join(f"{tot[w]:7d}" for w in WANT))
print("\nFP64 has no tcgen05 kind (tcgen05.mma: f16 / tf32 / f8f6f4 / i8 / mx*), so UTCMMA = 0 is by design: the FP64 tensor path of "
      "sm_100a is the warp-level DMMA.  UBLKCP = cp.async.bulk (TMA bulk copy engine, 1-D), LDGSTS = cp.async.")
