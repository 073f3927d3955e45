"""SASS instruction counts per kernel of the built library -> profiles/r02_sass_counts.txt
   python tools/sass_counts.py gpirt_b200/libgpirt_b200.so profiles/r02_sass_counts.txt"""
import collections
import re
import subprocess
import sys

lib, out = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
PAT = [("UTCIMMA", r"\bUTCIMMA\b"), ("LDTM/STTM", r"\b(LDTM|STTM)\b"), ("UTMALDG", r"\bUTMALDG\b"), ("UTMALDG.MULTICAST", r"UTMALDG\.\S*MULTICAST"),
       ("UBLKCP", r"\bUBLKCP\b"), ("SYNCS", r"\bSYNCS\b"), ("UCGABAR", r"\bUCGABAR"), ("DMMA", r"\bDMMA\b"), ("DFMA", r"\bDFMA\b"), ("LDGSTS", r"\bLDGSTS\b")]
chunks = re.split(r"\s+Function : \S+", sass)[1:]
lines = ["# SASS instruction counts per kernel of gpirt_b200/libgpirt_b200.so (sm_100a), round 2",
         "# command: python tools/sass_counts.py gpirt_b200/libgpirt_b200.so profiles/r02_sass_counts.txt  (cuobjdump -sass | per function)",
         "#   UTCIMMA (tcgen05.mma kind::i8)  LDTM/STTM (tcgen05.ld/st)  UTMALDG (TMA tensor load; .MULTICAST = cluster multicast)",
         "#   UBLKCP (bulk copy)  SYNCS (mbarrier)  UCGABAR (cluster barrier)  DMMA (FP64 tensor, mma.sync m8n8k4)  DFMA  LDGSTS (cp.async)"]
tot = collections.Counter()
for name, body in zip(names, chunks):
    short = re.sub(r"\(.*", "", re.sub(r"\((int|bool)\)", "", name)).replace("void ", "").replace("gpirt::", "").replace("(anonymous namespace)::", "")
    c = [(k, len(re.findall(p, body))) for k, p in PAT]
    c = [(k, v) for k, v in c if v]
    for k, v in c:
        tot[k] += v
    if c:
        lines.append("%-60s %s" % (short, "  ".join("%s=%d" % kv for kv in c)))
lines.append("TOTAL  " + "  ".join("%s=%d" % (k, tot[k]) for k, _ in PAT if tot[k]))
open(out, "w").write("\n".join(lines) + "\n")
print(lines[-1])
