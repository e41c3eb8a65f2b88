"""CPU restatement of the balanced base-256 digit scheme of the int8 contraction (csrc/dkg_ozaki.cu:
slice_rows_kernel + the diagonal cut-off of ozaki_kernel).  Checks on the host what the CUDA tests check on the
device -- digit range, exact representation, integer exactness -- and the claim the 7-diagonal cut-off rests on:
with zero-mean digits the dropped diagonals cost less than the rounding error of a plain fp64 GEMM."""
import numpy as np


def slice_rows(X, ns=7):
    """(digits [ns, rows, K] int64 in [-128, 127], power-of-two scale per row): x = scale * sum_s d_s 256^-s."""
    m = np.abs(X).max(axis=1)
    _, e = np.frexp(m)
    e7 = e - 7
    e7 = np.where(np.ldexp(m, -e7) > 127.0, e7 + 1, e7)  # carries may raise the leading digit by one: keep |x| <= 127
    V = np.rint(np.ldexp(X, (8 * (ns - 1) - e7)[:, None])).astype(np.int64)
    D = np.zeros((ns,) + X.shape, np.int64)
    for s in range(ns - 1, 0, -1):
        d = ((V + 128) & 255) - 128
        D[s] = d
        V = (V - d) >> 8
    D[0] = V
    return D, np.ldexp(1.0, e7)


def product(A, B, ns=7, ng=7):
    Da, sa = slice_rows(A, ns)
    Db, sb = slice_rows(B, ns)
    acc = np.zeros((A.shape[0], B.shape[0]), np.longdouble)
    for g in range(ng - 1, -1, -1):  # smallest weights first, as the epilogue folds them
        S = np.zeros((A.shape[0], B.shape[0]), np.int64)
        for i in range(ns):
            j = g - i
            if 0 <= j < ns:
                S += Da[i] @ Db[j].T
        assert np.abs(S).max() < 2**31  # the int32 accumulators of the tensor cores cannot overflow
        acc += S.astype(np.longdouble) * np.longdouble(256.0) ** (-g)
    return (acc * sa[:, None] * sb[None, :]).astype(np.float64)


def test_digits_are_int8_and_represent_the_rounded_value_exactly():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((64, 200)) * np.exp(rng.standard_normal((64, 1)) * 5)
    X[0, :4] = [127.0, 127.4999, -127.99, 1e-3]  # rows whose maximum lies above 127 after scaling give up one bit
    X[1, :3] = [128.0, -128.0, 0.5]
    X[2, :2] = [255.9, -255.9]
    D, s = slice_rows(X)
    assert D.min() >= -128 and D.max() <= 127
    rec = sum(D[i].astype(np.longdouble) * np.longdouble(256.0) ** (-i) for i in range(7)) * s[:, None]
    err = np.abs(rec - X) / np.abs(X).max(axis=1)[:, None]
    assert float(err.max()) <= 2.0**-54  # half a unit of the 7th digit (2^-49) over a row maximum scaled into [32, 127]


def test_small_integers_multiply_exactly():
    rng = np.random.default_rng(1)
    A = rng.integers(-(2**20), 2**20, (40, 96)).astype(np.float64)
    B = rng.integers(-(2**20), 2**20, (24, 96)).astype(np.float64)
    assert np.array_equal(product(A, B), A @ B.T)


def test_seven_diagonals_stay_below_the_rounding_error_of_an_fp64_gemm():
    rng = np.random.default_rng(400)
    M, N, K = 96, 128, 400
    A = rng.standard_normal((M, K)) * np.exp(rng.standard_normal((M, 1)) * 3)
    B = rng.standard_normal((N, K)) * np.exp(rng.standard_normal((N, 1)) * 3)
    truth = A.astype(np.longdouble) @ B.astype(np.longdouble).T
    bound = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * K
    err = {ng: float((np.abs(product(A, B, 7, ng) - truth) / bound).max()) for ng in (6, 7, 8, 13)}
    ref = float((np.abs(A @ B.T - truth) / bound).max())  # plain fp64 GEMM
    assert err[7] < ref, (err, ref)          # the shipped cut-off
    assert err[7] < 8 * err[13], err         # and within a small factor of keeping every diagonal
    assert err[6] > 20 * err[7], err         # one diagonal fewer would show
    assert err[7] < 4e-16
