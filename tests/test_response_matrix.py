"""Restates the reference's tests/testthat/test_response_matrix.R (its whole test suite) against the host mirror."""
import numpy as np
import pandas as pd
import pytest

from gpirt_b200 import ResponseMessage, as_response_matrix, is_response_matrix, response_matrix
from gpirt_b200.response_matrix import ResponseMatrix

NA = np.nan
x = [1, 0, 1, 1, 0, NA]
ex1 = np.array(x, dtype=float).reshape(3, 2, order="F")
ex2 = pd.DataFrame({"x1": x[:3], "x2": x[3:]})
codes01 = dict(yea=1, nay=0, missing=NA)
x3 = [1, -1, 2, 3, -1, NA]
ex3 = np.array(x3, dtype=float).reshape(3, 2, order="F")
ex4 = pd.DataFrame({"x1": x3[:3], "x2": x3[3:]})
codes_multi = dict(yea=[1, 2, 3], nay=-1, missing=NA)
ex5 = pd.DataFrame({"x": pd.Categorical(["Yea", "Nay", "Yea"]), "y": pd.Categorical(["Yea", "Nay", None])})
ex6 = pd.DataFrame({"x": pd.Categorical(["Yea", "Nay", "Yes"]), "y": pd.Categorical(["Yea", "Nay", None])})
codes_str = dict(yea="Yea", nay="Nay", missing=NA)


def value_set(rm):
    a = np.asarray(rm).ravel()
    return {("NA" if np.isnan(v) else float(v)) for v in a}


@pytest.mark.parametrize("data,codes", [(ex1, codes01), (ex2, codes01), (ex3, codes_multi), (ex4, codes_multi), (ex5, codes_str)])
def test_response_matrix_functions_properly(data, codes):       # test_response_matrix.R:53-63
    r = response_matrix(data, response_codes=codes)
    assert isinstance(r, ResponseMatrix) and r.r_class == "response_matrix"
    assert value_set(r) == {1.0, -1.0, "NA"}
    assert r.dtype == np.float64 and r.flags["F_CONTIGUOUS"]


def test_message_on_uncoded_value():                             # :64-66
    with pytest.warns(ResponseMessage, match="Yes were not given a response code"):
        response_matrix(ex6, response_codes=codes_str)


def test_error_on_list():                                        # :67
    with pytest.raises(TypeError, match="Conversion from lists"):
        response_matrix([1])


def test_is_response_matrix():                                   # :74-86
    all_true = ResponseMatrix(np.array([[1.0]]))
    values_wrong = ResponseMatrix(np.array([[6.0]]))
    class_false = np.array([[1.0]])
    matrix_false = ResponseMatrix(np.array([[1.0]]))[0]          # a classed non-matrix
    assert not is_response_matrix(class_false)
    assert not is_response_matrix(matrix_false)
    assert not is_response_matrix(values_wrong)
    assert is_response_matrix(all_true)


def test_as_response_matrix():                                   # :91-100
    r1 = response_matrix(ex1, response_codes=codes01)
    a = as_response_matrix(ex1, response_codes=codes01)
    assert np.array_equal(np.asarray(a), np.asarray(r1), equal_nan=True)
    b = as_response_matrix(r1, response_codes=codes01)
    assert b is r1                                               # identical(): untouched when already a response_matrix


def test_unanimous_items_are_dropped_with_message():             # R/response_matrix.R:80-88
    d = np.array([[1, 1, 0], [0, 1, 0], [1, 1, NA]], dtype=float)
    with pytest.warns(ResponseMessage, match="Items 2 and 3 were discarded as unanimous"):
        r = response_matrix(d, response_codes=codes01)
    assert r.shape == (3, 1) and np.array_equal(np.asarray(r)[:, 0], [1.0, -1.0, 1.0])


def test_default_codes_and_senate116_shape():
    import warnings
    import gpirt_b200
    codes, icpsr, rolls = gpirt_b200.senate116()
    assert codes.shape == (100, 428) and len(icpsr) == 100 and len(rolls) == 428
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        y = response_matrix(codes)                               # yea 1:3, nay 4:6, missing 0,7:9,NA
    assert y.shape == (100, 418), "10 unanimous roll calls dropped (SURVEY: 100 x 418)"
    assert any("discarded as unanimous" in str(m.message) for m in w)
    assert value_set(y) == {1.0, -1.0, "NA"} and abs(np.isnan(np.asarray(y)).mean() - 0.057) < 0.002
    # bit-exact ingest contract: the golden fixture's int8 coding is this matrix
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "senate116_100x418.npz"))
    assert np.array_equal(np.where(np.isnan(np.asarray(y)), 0, np.asarray(y)).astype(np.int8), g["y_int8"])
