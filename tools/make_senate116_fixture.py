"""Builds gpirt_b200/data/senate116_cast_codes.npz from the reference's raw Voteview CSVs
(/root/reference/data-raw/S116_votes.csv, S116_rollcalls.csv) following data-raw/senate116.R:5-9 (session-1 roll
calls only) and the vignette's reshape (vignettes/gpirt-vignette.Rmd:131-141: rows = icpsr ascending, columns =
rollnumber ascending, values = cast_code).  Run in the build container only; the .npz is committed because
/root/reference does not exist on the GPU box."""
import csv
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
with open(os.path.join(ref, "data-raw", "S116_rollcalls.csv"), newline="") as fh:
    session1 = {int(r["rollnumber"]) for r in csv.DictReader(fh) if int(r["session"]) == 1}
rows = []
with open(os.path.join(ref, "data-raw", "S116_votes.csv"), newline="") as fh:
    for r in csv.DictReader(fh):
        if int(r["rollnumber"]) in session1:
            rows.append((int(r["icpsr"]), int(r["rollnumber"]), int(r["cast_code"])))
icpsr = sorted({r[0] for r in rows})
rolls = sorted({r[1] for r in rows})
ii = {v: k for k, v in enumerate(icpsr)}
jj = {v: k for k, v in enumerate(rolls)}
codes = np.full((len(icpsr), len(rolls)), -1, dtype=np.int8)  # -1 = no row in the long table (NA after spread())
for a, b, c in rows:
    codes[ii[a], jj[b]] = c
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpirt_b200", "data", "senate116_cast_codes.npz")
np.savez_compressed(out, cast_code=codes, icpsr=np.array(icpsr, dtype=np.int32), rollnumber=np.array(rolls, dtype=np.int32))
print(codes.shape, "long rows", len(rows), "absent cells", int((codes < 0).sum()))
