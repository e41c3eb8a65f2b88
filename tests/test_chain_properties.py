"""Properties the tile-first filter and the champion probe rest on, restated in numpy (CPU only; the kernels are
pinned on the GPU by test_gpu_filter32.py):

1. ANY member lines of a set may serve as the vertices U / V of the second-level chain P-U-T-V-Q (the probe -- tile
   champions or a line sample -- only decides how tight the chain is): no vertex of the upper envelope is ever
   below it, so the hull found among the survivors is the hull of the whole set.
2. The tile cull is exact: a tile whose intercept maximum does not exceed the chain at either end of the tile's
   slope range holds no survivor (the chords are linear in the slope), and a chain SIDE that is dead for the tile
   keeps no line of it.
"""
import numpy as np
import pytest
import torch

from oracle import discretekg as odk


def _chords(verts):
    """[(c, m)] of the chords through consecutive vertices (z, a) with strictly increasing z (concave or not)."""
    out = []
    for (z0, a0), (z1, a1) in zip(verts[:-1], verts[1:]):
        if z1 > z0:
            m = (a1 - a0) / (z1 - z0)
            out.append((a0 - m * z0, m))
    return out


def _chain5(z, a, pick):
    """P-U-T-V-Q with U / V = the farthest line of `pick` above the chord P-T / T-Q (as chain5_kernel builds it)."""
    iP, iQ, iT = int(np.argmin(z)), int(np.argmax(z)), int(np.argmax(a))
    P, Q, T = (z[iP], a[iP]), (z[iQ], a[iQ]), (z[iT], a[iT])
    verts = [P]
    for lo, hi in ((P, T), (T, Q)):
        if hi[0] > lo[0]:
            m = (hi[1] - lo[1]) / (hi[0] - lo[0])
            cand = [k for k in pick if lo[0] < z[k] < hi[0] and a[k] > lo[1] + m * (z[k] - lo[0])]
            if cand:
                k = max(cand, key=lambda k: a[k] - (lo[1] + m * (z[k] - lo[0])))
                verts.append((z[k], a[k]))
        if hi[0] > verts[-1][0] or hi is T:
            verts.append(hi)
    # strictly increasing slopes only
    v2 = [verts[0]]
    for v in verts[1:]:
        if v[0] > v2[-1][0]:
            v2.append(v)
    return v2


def _survivors(z, a, verts):
    t = np.full(z.shape, np.inf)
    for c, m in _chords(verts):
        t = np.minimum(t, c + m * z)
    return a > t - 1e-12 * (np.abs(a).max() + 1.0)  # (the kernels keep lines within 128 ulp of the chain)


@pytest.mark.parametrize("seed", range(12))
def test_any_member_lines_as_second_level_vertices_keep_the_whole_hull(seed):
    rng = np.random.default_rng(seed)
    n = 1024
    x = np.sort(rng.random(n))  # a smooth-ish "posterior" along a 1-d ordering + noise: many near-hull lines
    a = np.sin(3 * x + rng.random()) + 0.05 * rng.standard_normal(n)
    z = np.cos(2 * x + rng.random()) * (0.5 + rng.random()) + 0.05 * rng.standard_normal(n)
    idx, _ = odk.epigraph_indices(torch.from_numpy(a), torch.from_numpy(z))  # the reference's march over ALL lines
    hull = set(int(i) for i in idx)
    tile = 128
    champions = [int(t * tile + np.argmax(a[t * tile:(t + 1) * tile])) for t in range(n // tile)]
    sample = [k for k in range(n) if k % 256 < 16]
    arbitrary = list(rng.choice(n, 40, replace=False))
    counts = {}
    for name, pick in (("champions", champions), ("sample", sample), ("arbitrary", arbitrary), ("none", [])):
        verts = _chain5(z, a, pick)
        keep = _survivors(z, a, verts)
        assert hull <= set(np.nonzero(keep)[0].tolist()), name
        counts[name] = int(keep.sum())
    assert counts["champions"] <= counts["none"] and counts["sample"] <= counts["none"]


@pytest.mark.parametrize("seed", range(6))
def test_tile_cull_and_side_masks_are_exact(seed):
    rng = np.random.default_rng(100 + seed)
    n, tile = 2048, 128
    x = np.sort(rng.random(n))
    a = np.sin(4 * x) + 0.02 * rng.standard_normal(n)
    z = np.cos(3 * x) + 0.02 * rng.standard_normal(n)  # neighbouring lines have neighbouring slopes (Morton order)
    champions = [int(t * tile + np.argmax(a[t * tile:(t + 1) * tile])) for t in range(n // tile)]
    verts = _chain5(z, a, champions)
    ch = _chords(verts)
    iT = max(range(len(verts)), key=lambda k: verts[k][1])
    left, right = ch[:iT], ch[iT:]
    for t in range(n // tile):
        zs, as_ = z[t * tile:(t + 1) * tile], a[t * tile:(t + 1) * tile]
        zlo, zhi, am = zs.min(), zs.max(), as_.max()

        def dead(side):
            return all(am <= min(c + m * zlo, c + m * zhi) for c, m in side)

        def passes(side):
            return np.zeros(tile, bool) if not side else np.any([as_ > c + m * zs for c, m in side], axis=0)

        if dead(left):
            assert not passes(left).any()
        if dead(right):
            assert not passes(right).any()
        if dead(left) and dead(right):
            assert not (passes(left) | passes(right)).any()
