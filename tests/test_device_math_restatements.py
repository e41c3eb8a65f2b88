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
