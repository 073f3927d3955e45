"""Bundled example data: U.S. Senate roll calls, 116th Congress, first session (reference data/senate116.rda,
R/senate116.R; built by tools/make_senate116_fixture.py from the reference's data-raw CSVs)."""
import os

import numpy as np


def senate116(wide=True):
    """Returns the 100 x 428 matrix of Voteview cast codes (rows = icpsr ascending, columns = rollnumber ascending),
    i.e. the vignette's `responses` after spread(); pass it to response_matrix() / gpirtMCMC() with the default
    vote codes.  wide=False returns the long table (icpsr, rollnumber, cast_code) like the reference's data frame."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "senate116_cast_codes.npz"))
    codes = z["cast_code"].astype(np.float64)
    if wide:
        from ..response_matrix import ResponseMatrix  # only for the dimnames carrier
        out = codes.view(np.ndarray)
        return out, [str(i) for i in z["icpsr"]], [str(r) for r in z["rollnumber"]]
    ii, jj = np.meshgrid(z["icpsr"], z["rollnumber"], indexing="ij")
    return np.column_stack([ii.ravel(), jj.ravel(), codes.ravel()])
