"""numpy restatements of the branch-free device helpers of dkg_common.cuh / dkg_coupled.cu, to pin the
ALGORITHMS' accuracy on the CPU (the kernels themselves are pinned against the oracle on the GPU):
exp_nonpos (Cody-Waite reduction + degree-13 Taylor), sqrt_pos (reciprocal-square-root seed, one
third-order step, one correction) and the reciprocal + residual-step quotient of the coupled slopes."""
import math

import numpy as np

TAYLOR = [1.0 / math.factorial(k) for k in range(13, 1, -1)]  # 1/13! ... 1/2!


def exp_nonpos(x):
    x = np.asarray(x, dtype=np.float64)
    xc = np.where(x < -708.0, -708.0, x)
    t = xc * 1.4426950408889634074 + 6755399441055744.0
    kd = t - 6755399441055744.0
    r = xc - kd * 6.93147180369123816490e-01
    r = r - kd * 1.90821492927058770002e-10
    p = np.full_like(r, TAYLOR[0])
    for c in TAYLOR[1:]:
        p = p * r + c
    p = p * r + 1.0
    p = p * r + 1.0
    with np.errstate(invalid="ignore"):
        k = np.where(np.isfinite(kd), kd, 0.0).astype(np.int64)
    return np.ldexp(p, k)


def sqrt_pos(x, seed_bits=22):
    x = np.asarray(x, dtype=np.float64)
    y = 1.0 / np.sqrt(x)
    y = y * (1.0 + 2.0 ** -seed_bits * np.sign(np.sin(np.arange(x.size) + 0.5)))  # a 2^-22 seed error
    e = 1.0 - x * (y * y)
    y1 = y + (0.5 + 0.375 * e) * (y * e)
    g = x * y1
    return g + (x - g * g) * (0.5 * y1)


def ulps(got, want):
    return np.abs(got - want) / np.spacing(np.abs(want))


def test_exp_nonpos_is_faithful():
    rng = np.random.default_rng(0)
    x = -np.concatenate([rng.uniform(0, 40, 200000), rng.uniform(0, 700, 50000), 10.0 ** rng.uniform(-300, 0, 20000), [0.0]])
    got, want = exp_nonpos(x), np.exp(x)
    assert ulps(got, want).max() <= 2.0  # (numpy has no fma: the device version is at least as tight)
    assert exp_nonpos(np.array([-1e6]))[0] < 1e-300  # anything below -708 is returned as ~3e-308
    assert np.isnan(exp_nonpos(np.array([np.nan]))[0])


def test_sqrt_pos_is_faithful():
    rng = np.random.default_rng(1)
    x = np.concatenate([10.0 ** rng.uniform(-30, 30, 200000), rng.uniform(0.5, 2.0, 100000), [1e-30, 1.0, 4.0]])
    got, want = sqrt_pos(x), np.sqrt(x)
    assert ulps(got, want).max() <= 1.0


def test_reciprocal_residual_quotient_matches_division():
    """q = s r;  q' = fma(fma(-q, sd, s), r, q) with r = RN(1 / sd): exact rational arithmetic emulates the
    two fused multiply-adds (one rounding each)."""
    from fractions import Fraction

    rng = np.random.default_rng(2)
    n = 20000
    s = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 8, n)
    sd = 10.0 ** rng.uniform(-6, 6, n)
    exact = 0
    for sv, dv in zip(s.tolist(), sd.tolist()):
        r = 1.0 / dv
        q = sv * r
        resid = float(Fraction(sv) - Fraction(q) * Fraction(dv))
        got = float(Fraction(q) + Fraction(resid) * Fraction(r))
        want = sv / dv
        assert abs(got - want) <= math.ulp(want)
        exact += got == want
    assert exact >= n - 2  # correctly rounded (Markstein), up to the rare all-ones-significand divisors


def order_key(s):
    """dkg_coupled.cu::order_key: high word of the double, sign-magnitude -> two's complement."""
    hi = (np.asarray(s, dtype=np.float64).view(np.int64) >> 32).astype(np.int32)
    return hi ^ ((hi >> 31) & np.int32(0x7FFFFFFF))


def test_coarse_keys_are_monotone_so_the_exact_extreme_lies_in_a_segment_with_the_extreme_key():
    """The row statistics keep only a coarse, MONOTONE image of every value per segment / tile -- the float rounding in
    zfinish_tiled_kernel, the integer order key in coupled_range_kernel -- and re-read exactly just the segments that
    attain the extreme image.  Monotone (non-strict) is all that needs: x <= y  =>  image(x) <= image(y)."""
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.standard_normal(20000) * np.exp(rng.standard_normal(20000) * 8),
                        [0.0, 1e-310, -1e-310, 1e308, -1e308, np.inf, -np.inf, 1.0, np.nextafter(1.0, 2.0), -1.0, np.nextafter(-1.0, -2.0)]])
    x = np.sort(x)
    k = order_key(x).astype(np.int64)
    assert np.all(k[1:] >= k[:-1])
    with np.errstate(over="ignore"):
        f = x.astype(np.float32)
    assert np.all(f[1:] >= f[:-1])
    # segments of 512 values in arbitrary order: the exact minimum / maximum sit in a segment that attains the extreme key
    y = rng.permutation(x[np.isfinite(x)])[: 512 * 39].reshape(39, 512)
    for image in (lambda v: order_key(v).astype(np.int64), lambda v: v.astype(np.float32)):
        im = image(y)
        lo_seg = np.nonzero(im.min(axis=1) == im.min())[0]
        hi_seg = np.nonzero(im.max(axis=1) == im.max())[0]
        assert y.min() == y[lo_seg].min() and y.max() == y[hi_seg].max()
