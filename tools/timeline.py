"""Prints the timeline of one sweep from a GPIRT_TRACE file (segments of all streams, sorted by start).
   GPIRT_TRACE=/tmp/t.txt python bench.py ... ; python tools/timeline.py /tmp/t.txt [which | graph]
   which: index of the ESS segment that starts the sweep in the eager trace (default -3); graph: the captured sweep
   (globaltimer stamps of the last graph replay of the largest sampler)."""
import sys
NAMES = ["fill_z", "lz_gemm", "ess", "kstar", "trsm", "fstar_gemm", "fstar_draw", "theta_prep", "theta_gemm", "allreduce", "theta_draw",
         "beta", "kbuild", "chol", "trtri"]
blocks = []
for l in open(sys.argv[1]):
    if l.startswith("#"):
        blocks.append((l.strip(), []))
    elif l.strip() and blocks:
        t, a, b = l.split()
        blocks[-1][1].append((int(t), float(a), float(b)))
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
cand = [b for b in blocks if ("graph sweep" in b[0]) == graph]
head, rows = max(cand, key=lambda b: len(b[1]))
print(head)
if graph:
    t0, t1 = min(a for _, a, _ in rows), max(b for _, _, b in rows)
    print("captured sweep: %.3f ms from the first to the last stamp" % (t1 - t0))
else:
    ess = sorted(a for t, a, b in rows if t == 2)
    which = int(sys.argv[2]) if len(sys.argv) > 2 else -3
    t0, t1 = ess[which], ess[which + 1]
    print("sweep of %.3f ms (ESS start to next ESS start)" % (t1 - t0))
for t, a, b in sorted(rows, key=lambda r: r[1]):
    if t0 <= a < t1 or graph:
        print("  %-11s %8.3f -> %8.3f  (%7.3f ms)" % (NAMES[t], a - t0, b - t0, b - a))
