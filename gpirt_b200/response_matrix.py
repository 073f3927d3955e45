"""Response coding — the host-side step in front of the sampler.

Mirror of the reference's R/response_matrix.R:51-127 (response_matrix / is.response_matrix / as.response_matrix):
same argument meaning, same messages, same errors, same output contract — an n x m float64 matrix holding exactly
{+1, -1, NaN(NA)}, unanimous items dropped — so the tests in tests/test_response_matrix.py read like the reference's
tests/testthat/test_response_matrix.R.  The sampler consumes the matrix as column-major float64 (what R hands the
native routine, src/RcppExports.cpp:20) and ingests it on the device as int8 {+1,-1,0}.
"""
import math
import warnings

import numpy as np

DEFAULT_CODES = dict(yea=[1, 2, 3], nay=[4, 5, 6], missing=[0, 7, 8, 9, None])


class ResponseMessage(UserWarning):
    """Stands in for R's message(): informational, not an error."""


class ResponseMatrix(np.ndarray):
    """float64 matrix of {+1,-1,NaN} carrying the reference's S3 class tag 'response_matrix' plus dim names."""

    r_class = "response_matrix"

    def __new__(cls, arr, rownames=None, colnames=None):
        obj = np.asfortranarray(np.asarray(arr, dtype=np.float64)).view(cls)
        obj.rownames = rownames
        obj.colnames = colnames
        obj._values_ok = None
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.rownames = getattr(obj, "rownames", None)
        self.colnames = getattr(obj, "colnames", None)
        self._values_ok = None   # cache of the {NA,-1,1} scan; views and slices start unchecked


def _is_na(v):
    if v is None:
        return True
    try:
        return isinstance(v, float) and math.isnan(v) or (hasattr(v, "dtype") and np.issubdtype(v.dtype, np.floating) and np.isnan(v))
    except TypeError:
        return False


def _key(v):
    """R's %in% compares after coercion: 1L, 1.0 and (against character data) "1" all match."""
    if _is_na(v):
        return ("na",)
    if isinstance(v, (bool, np.bool_)):
        return ("num", float(v))
    if isinstance(v, (int, float, np.integer, np.floating)):
        return ("num", float(v))
    s = str(v)
    try:
        return ("num", float(s))
    except ValueError:
        return ("str", s)


def _printc(words):
    """R/response_matrix.R:3-9"""
    words = [str(w) for w in words]
    n = len(words)
    if n == 1:
        return words[0]
    words[-1] = "and " + words[-1]
    return " ".join(words) if n == 2 else ", ".join(words)


def _fmt(v):
    if _is_na(v):
        return "NA"
    if isinstance(v, (float, np.floating)) and float(v).is_integer():
        return str(int(v))
    return str(v)


def _as_object_matrix(data):
    """as.matrix(): returns (object ndarray n x m, rownames, colnames)."""
    try:
        import pandas as pd
    except Exception:  # pragma: no cover
        pd = None
    if pd is not None and isinstance(data, pd.DataFrame):
        rn = [str(x) for x in data.index] if not isinstance(data.index, pd.RangeIndex) else None
        cn = [str(c) for c in data.columns]
        cols = []
        for c in data.columns:
            col = data[c]
            vals = col.astype(object).to_numpy()
            cols.append(np.array([None if (v is None or (isinstance(v, float) and math.isnan(v)) or v is pd.NA) else v for v in vals], dtype=object))
        return np.stack(cols, axis=1) if cols else np.empty((len(data), 0), dtype=object), rn, cn
    arr = np.asarray(data)
    if arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    if arr.ndim != 2:
        raise ValueError("response data must be two-dimensional")
    return arr.astype(object), getattr(data, "rownames", None), getattr(data, "colnames", None)


def response_matrix(data, response_codes=None):
    """R/response_matrix.R:51-99.  `response_codes` maps 'yea' -> +1, 'nay' -> -1, 'missing' -> NA."""
    if isinstance(data, (list, tuple, dict)):
        # R: is.list(data) & !is.data.frame(data)
        raise TypeError("Conversion from lists to response_matrix objects is currently unsupported.")
    codes = {k: list(np.atleast_1d(np.asarray(v, dtype=object))) for k, v in (response_codes or DEFAULT_CODES).items()}
    for k in ("yea", "nay", "missing"):
        codes.setdefault(k, [])
    raw, rnames, cnames = _as_object_matrix(data)
    n, m = raw.shape
    keys = np.empty((n, m), dtype=object)
    for j in range(m):
        for i in range(n):
            keys[i, j] = _key(raw[i, j])
    yea = {_key(v) for v in codes["yea"]}
    nay = {_key(v) for v in codes["nay"]}
    mis = {_key(v) for v in codes["missing"]}
    known = yea | nay | mis
    omitted = []
    for j in range(m):  # setdiff() keeps first-appearance order, column-major
        for i in range(n):
            if keys[i, j] not in known and keys[i, j] not in [_key(o) for o in omitted]:
                omitted.append(raw[i, j])
    if omitted:
        mis |= {_key(o) for o in omitted}
        warnings.warn("Responses with value " + _printc([_fmt(o) for o in omitted]) + " were not given a response code "
                      "and will be treated as missing.", ResponseMessage, stacklevel=2)
    result = np.full((n, m), np.nan, dtype=np.float64)
    # assignment order as the reference: yea, then nay, then missing (later rules win on overlapping codes)
    for rule, val in ((yea, 1.0), (nay, -1.0), (mis, np.nan)):
        for j in range(m):
            for i in range(n):
                if keys[i, j] in rule:
                    result[i, j] = val
    # guard against unanimity: length(unique(na.omit(x))) == 1
    unanimous = np.array([np.unique(result[~np.isnan(result[:, j]), j]).size == 1 for j in range(m)], dtype=bool)
    kept = result[:, ~unanimous]
    if unanimous.any():
        which = [str(j + 1) if cnames is None else str(cnames[j]) for j in np.flatnonzero(unanimous)]
        nu = int(unanimous.sum())
        warnings.warn("Item" + ("s " if nu > 1 else " ") + _printc(which) + (" were" if nu > 1 else " was") +
                      " discarded as unanimous.", ResponseMessage, stacklevel=2)
    kept_names = None if cnames is None else [c for c, u in zip(cnames, unanimous) if not u]
    out = ResponseMatrix(kept, rownames=rnames, colnames=kept_names)
    out._values_ok = True
    return out


def is_response_matrix(x):
    """R/response_matrix.R:108-114: class tag, is a matrix, and values within {NA,-1,1}."""
    if not isinstance(x, ResponseMatrix):
        return False
    if x.ndim != 2:
        return False
    if x._values_ok is None:   # one O(nm) scan per object (arrays are not expected to be mutated afterwards)
        a = np.asarray(x)
        ok = True
        step = max(1, (1 << 22) // max(1, a.shape[0]))
        for j0 in range(0, a.shape[1], step):
            blk = a[:, j0:j0 + step]
            if not np.all(np.isnan(blk) | (np.abs(blk) == 1.0)):
                ok = False
                break
        x._values_ok = ok
    return bool(x._values_ok)


def as_response_matrix(x, response_codes=None):
    """R/response_matrix.R:119-127"""
    if not is_response_matrix(x):
        x = response_matrix(x, response_codes)
    return x
