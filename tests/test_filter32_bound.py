"""The float chord filter must never drop a line the float64 filter keeps.  This restates the rounding
rule of ``chord32`` (decoupled-kg_b200/csrc/dkg_emax.cu) in numpy and checks the inclusion on random and
adversarial (near-chain, large-offset, tiny-slope) line sets.  CPU only: it pins the error analysis, the
GPU test ``test_gpu_filter32.py`` pins the kernels."""
import numpy as np
import pytest

EPS128 = 128 * 2.0 ** -52


def chain_fp64(z, a):
    """(c1, m1, c2, m2) of the chain P -> T -> Q with the kernel's 128-ulp slack (chain_params)."""
    iP, iQ, iT = int(np.argmin(z)), int(np.argmax(z)), int(np.argmax(a))
    zP, zQ, zT, aP, aQ, aT = z[iP], z[iQ], z[iT], a[iP], a[iQ], a[iT]
    c1 = c2 = np.inf
    m1 = m2 = 0.0
    if zT > zP:
        m1 = (aT - aP) / (zT - zP)
        c1 = aP - m1 * zP - EPS128 * (abs(aT) + abs(aP) + abs(m1) * max(abs(zP), abs(zT)))
    if zQ > zT:
        m2 = (aQ - aT) / (zQ - zT)
        c2 = aT - m2 * zT - EPS128 * (abs(aT) + abs(aQ) + abs(m2) * max(abs(zQ), abs(zT)))
    amag = max(abs(aT), abs(aP), abs(aQ))
    zmag = max(abs(z.min()), abs(z.max()))
    return (c1, m1, c2, m2), amag, zmag


def chord32(c, m, amag, zmag):
    """float image (m32, c32) of one chord: c lowered by 2^-21 of the magnitudes, rounded DOWN."""
    if np.isinf(c):
        return np.float32(0.0), np.float32(np.inf)
    g = abs(c) + abs(m) * zmag + amag
    ok = 1e-30 < g < 1e30 and abs(m) < 1e30 and zmag < 1e30 and (zmag + abs(m)) <= g * 1e37
    if not ok:
        return np.float32(0.0), np.float32(-np.inf)
    cs = c - 2.0 ** -21 * g
    c32 = np.float32(cs)
    if float(c32) > cs:  # round toward -inf
        c32 = np.nextafter(c32, np.float32(-np.inf))
    return np.float32(m), c32


def fma32(m32, z32, c32):
    """float32 fma: the product of two floats is exact in float64; one rounding to float32 follows
    (double rounding through float64 can only matter at a 2^-53 tie, far inside the slack)."""
    return (m32.astype(np.float64) * z32.astype(np.float64) + c32.astype(np.float64)).astype(np.float32)


def survivors(z, a):
    (c1, m1, c2, m2), amag, zmag = chain_fp64(z, a)
    t1 = m1 * z + c1 if np.isfinite(c1) else np.full_like(z, np.inf)
    t2 = m2 * z + c2 if np.isfinite(c2) else np.full_like(z, np.inf)
    keep64 = (a > t1) | (a > t2)
    z32, a32 = z.astype(np.float32), a.astype(np.float32)
    m1f, c1f = chord32(c1, m1, amag, zmag)
    m2f, c2f = chord32(c2, m2, amag, zmag)
    with np.errstate(invalid="ignore", over="ignore"):
        u1 = fma32(np.full_like(z32, m1f), z32, np.full_like(z32, c1f))
        u2 = fma32(np.full_like(z32, m2f), z32, np.full_like(z32, c2f))
        keep32 = (a32 > u1) | (a32 > u2)
    return keep64, keep32


def _sets(rng):
    n = 4000
    # GP-like: smooth curve in the dual plane plus noise of very different sizes
    for noise in (1e-1, 1e-4, 1e-8, 1e-12):
        t = np.sort(rng.uniform(-1, 1, n))
        z = t + noise * rng.standard_normal(n)
        a = 1.0 - t * t + noise * rng.standard_normal(n)
        yield z, a
    # large common offset in the intercepts (float rounding of a is ~1e-7 * offset)
    t = rng.uniform(-1, 1, n)
    yield t, 1e4 + (1 - t * t) * 1e-3 + 1e-9 * rng.standard_normal(n)
    # tiny slopes, huge chord gradients
    t = rng.uniform(-1, 1, n)
    yield 1e-7 * t, 5.0 - t * t + 1e-10 * rng.standard_normal(n)
    # lines exactly ON the chords and within a few ulps of them
    z = np.linspace(-1.0, 1.0, n)
    a = 1.0 - np.abs(z)
    a[1:-1:7] = np.nextafter(a[1:-1:7], np.inf)
    a[2:-1:7] = np.nextafter(a[2:-1:7], -np.inf)
    yield z, a
    # negative intercepts, asymmetric ranges
    t = rng.uniform(0, 3, n)
    yield t - 2.5, -7.0 - (t - 1) ** 2 + 1e-6 * rng.standard_normal(n)


@pytest.mark.parametrize("seed", range(5))
def test_float_filter_keeps_every_float64_survivor(seed):
    rng = np.random.default_rng(seed)
    for z, a in _sets(rng):
        keep64, keep32 = survivors(z, a)
        assert not np.any(keep64 & ~keep32)
        # and it is still a filter: it does not keep (much) more than the exact one
        assert keep32.sum() <= keep64.sum() + 0.02 * len(z) + 8


def test_out_of_range_magnitudes_keep_everything():
    rng = np.random.default_rng(0)
    t = rng.uniform(-1, 1, 500)
    for scale in (1e35, 1e-35):
        z, a = t.copy(), scale * (1 - t * t)
        keep64, keep32 = survivors(z, a)
        assert not np.any(keep64 & ~keep32)
        assert keep32.all()  # the rule falls back to "every line survives" (exact overflow path)
