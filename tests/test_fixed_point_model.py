"""CPU model of the fixed-point product the tensor-core kernel evaluates (gpirt_b200/csrc/dgemm_i8.cu), in exact Python
integers / numpy int64: pins the arithmetic of the scheme itself — digit decomposition, the 36 plane pairs, the level
recombination and the stated error bound — independently of any GPU.  The GPU kernel is tested against the same bound in
tests/test_gpu_parity.py::test_dgemm_i8_fixed_point_tensor_core_product."""
import numpy as np
import pytest

S = 8   # digit planes per operand


def digits(x_row):
    """balanced base-128 digits of round(x 2^(55-e)), e = ilogb(max|x|) + 1; returns (d[S, len], scale = 2^(e-6))"""
    mx = np.max(np.abs(x_row))
    e = int(np.floor(np.log2(mx))) + 1 if mx > 0 else 0
    X = np.rint(x_row * 2.0 ** (55 - e)).astype(np.int64)
    assert np.all(np.abs(X) <= 2 ** 55)
    d = np.zeros((S, x_row.size), dtype=np.int64)
    for s in range(S - 1, 0, -1):
        dd = ((X + 64) & 127) - 64
        d[s] = dd
        X = (X - dd) >> 7
    d[0] = X
    assert np.all(np.abs(d) <= 64), "every digit fits a signed byte"
    return d, 2.0 ** (e - 6)


def recompose(d, scale):
    return scale * sum(d[s].astype(np.float64) * 128.0 ** -s for s in range(S))


def fixed_point_matmul(A, B):
    """C = A @ B evaluated as the kernel does: per-row digit planes of A, per-column planes of B, exact integer plane
    products for levels s + t <= 7, Horner over the levels in FP64, two power-of-two scales"""
    M, K = A.shape
    N = B.shape[1]
    da, sa = zip(*(digits(A[i]) for i in range(M)))
    db, sb = zip(*(digits(B[:, j]) for j in range(N)))
    C = np.zeros((M, N))
    for i in range(M):
        for j in range(N):
            level = np.zeros(S, dtype=np.int64)
            for s in range(S):
                for t in range(S - s):
                    acc = int(np.dot(da[i][s], db[j][t]))
                    assert abs(acc) < 2 ** 31, "int32 accumulator of the tensor core"
                    level[s + t] += acc
            assert np.all(np.abs(level) < 2 ** 31)
            x = float(level[S - 1])
            for lv in range(S - 2, -1, -1):
                x = x * 0.0078125 + float(level[lv])
            C[i, j] = x * sa[i] * sb[j]
    return C


def test_digit_planes_reproduce_the_operand_to_56_bits():
    rs = np.random.RandomState(0)
    for scale in (1e-8, 1.0, 37.5, 1e12):
        x = rs.randn(257) * scale
        x[3] = 0.0
        d, sc = digits(x)
        assert np.max(np.abs(recompose(d, sc) - x)) <= 2.0 ** -55 * np.max(np.abs(x)) * 2
    d, sc = digits(np.zeros(5))
    assert not d.any()


@pytest.mark.parametrize("M,N,K", [(7, 5, 64), (4, 6, 1000), (3, 3, 4096)])
def test_fixed_point_product_obeys_the_stated_bound(M, N, K):
    rs = np.random.RandomState(K)
    A = rs.randn(M, K) * np.exp(2.0 * rs.randn(M, 1))
    B = rs.randn(K, N) * np.exp(2.0 * rs.randn(1, N))
    C = fixed_point_matmul(A, B)
    ref = (A.astype(np.longdouble) @ B.astype(np.longdouble)).astype(np.float64)
    bound = K * 2.0 ** -51 * np.abs(A).max(axis=1)[:, None] * np.abs(B).max(axis=0)[None, :]
    assert np.all(np.abs(C - ref) <= bound)
    # and it is as good as an FP64 dot product when rows are evenly scaled
    A2, B2 = rs.randn(M, K), rs.randn(K, N)
    C2 = fixed_point_matmul(A2, B2)
    ref2 = (A2.astype(np.longdouble) @ B2.astype(np.longdouble)).astype(np.float64)
    assert np.max(np.abs(C2 - ref2)) <= 64 * 2.0 ** -53 * np.max(np.abs(A2) @ np.abs(B2))


def test_worst_case_accumulators_fit_int32_below_k_65536():
    # level l holds (l + 1) plane pairs of |digit| <= 64 over K terms
    assert 8 * 64 * 64 * 65535 < 2 ** 31   # the library refuses K >= 65536


def test_fixed_scale_digits_of_normal_draws_and_cholesky_rows():
    """Z uses one scale 2^4 for every column (|z| < 8.7 for a 53-bit Box-Muller draw), L one scale 2^1 for every row
    (|L_ik| <= sqrt(1.001)): digits must still fit a signed byte"""
    zmax = np.sqrt(-2.0 * np.log(2.0 ** -54))
    assert zmax < 16.0
    for v, e in ((zmax, 4), (-zmax, 4), (np.sqrt(1.001), 1), (-np.sqrt(1.001), 1)):
        X = int(np.rint(v * 2.0 ** (55 - e)))
        ds = []
        for _ in range(S - 1):
            dd = ((X + 64) & 127) - 64
            ds.append(dd)
            X = (X - dd) >> 7
        ds.append(X)
        assert all(abs(d) <= 64 for d in ds)
