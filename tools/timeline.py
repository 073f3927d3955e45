"""Prints the timeline of the last complete sweep from a GPIRT_TRACE file (segments of all streams, sorted by start):
   GPIRT_TRACE=/tmp/t.txt python bench.py ... ; python tools/timeline.py /tmp/t.txt"""
import sys
NAMES = ["fill_z", "lz_gemm", "ess", "kstar", "trsm", "fstar_gemm", "fstar_draw", "theta_prep", "theta_gemm", "allreduce", "theta_draw",
         "beta", "kbuild", "chol", "trtri"]
blocks = [[]]
for l in open(sys.argv[1]):
    if l.startswith("#"):
        blocks.append([])
    elif l.strip():
        t, a, b = l.split()
        blocks[-1].append((int(t), float(a), float(b)))
rows = max(blocks, key=len)   # the sampler with the most timed segments (the benchmarked one)
ess = sorted(a for t, a, b in rows if t == 2)
which = int(sys.argv[2]) if len(sys.argv) > 2 else -3
t0, t1 = ess[which], ess[which + 1]
print("sweep of %.3f ms (ESS start to next ESS start)" % (t1 - t0))
for t, a, b in sorted(rows, key=lambda r: r[1]):
    if t0 <= a < t1:
        print("  %-11s %8.3f -> %8.3f  (%7.3f ms)" % (NAMES[t], a - t0, b - t0, b - a))
