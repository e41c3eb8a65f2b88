"""The int8 tensor-core contraction (csrc/dkg_ozaki.cu) against exact / extended-precision products.

Digits are exact, so with integer inputs small enough for one digit the product must be bit-exact;
with real inputs the error must stay at the level of an fp64 dot product (relative to
|row|max * |col|max * K, the bound of the fixed-point scheme)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mm(A, Bt, **kw):
    from decoupledbo_b200 import _native

    return _native.int8_matmul(A, Bt, **kw).cpu()


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 128), (128, 256, 416), (300, 200, 100), (1, 1, 1), (257, 129, 800)])
def test_small_integers_are_exact(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randint(-100, 100, (M, K), generator=g).double()
    B = torch.randint(-100, 100, (N, K), generator=g).double()
    want = A @ B.T
    got = _mm(A, B)
    assert torch.equal(got, want)


@pytest.mark.parametrize("nd", [1, 2, 3])
def test_digit_planes_and_signs(nd):
    # values with exactly `nd` base-256 digits below the leading one, both signs
    g = torch.Generator().manual_seed(nd)
    A = torch.randint(-(2 ** (8 * nd - 2)), 2 ** (8 * nd - 2), (128, 64), generator=g).double()
    B = torch.randint(-(2 ** (8 * nd - 2)), 2 ** (8 * nd - 2), (128, 64), generator=g).double()
    want = A @ B.T  # exact in fp64 while 2 * (8 nd - 2) + 6 <= 53
    got = _mm(A, B, n_digits=nd + 1, n_diagonals=2 * nd + 1)
    assert torch.equal(got, want)


@pytest.mark.parametrize("M,N,K", [(256, 384, 400), (128, 128, 1000)])
def test_real_inputs_match_extended_precision(M, N, K):
    rng = np.random.default_rng(K)
    A = rng.standard_normal((M, K)) * np.exp(rng.standard_normal((M, 1)) * 3)
    B = rng.standard_normal((N, K)) * np.exp(rng.standard_normal((N, 1)) * 3)
    truth = (A.astype(np.longdouble) @ B.astype(np.longdouble).T).astype(np.float64)
    got = _mm(torch.from_numpy(A), torch.from_numpy(B)).numpy()
    bound = np.abs(A).max(1)[:, None] * np.abs(B).max(1)[None, :] * K
    err = np.abs(got - truth) / bound
    ref = np.abs(A @ B.T - truth) / bound  # plain fp64 GEMM on the host
    assert err.max() < 4e-16, (err.max(), ref.max())


def test_cta_pair_kernel_is_exact_too():
    """DKG_OZ_PAIR=1 routes the default digit configuration through the cluster-of-2 kernel
    (tcgen05.mma.cta_group::2); the mode is latched per process, hence the subprocess."""
    import os
    import subprocess
    import sys

    code = (
        "import sys, torch; sys.path.insert(0, %r); from decoupledbo_b200 import _native;"
        "g = torch.Generator().manual_seed(5);"
        "A = torch.randint(-100, 100, (300, 416), generator=g).double();"
        "B = torch.randint(-100, 100, (200, 416), generator=g).double();"
        "assert torch.equal(_native.int8_matmul(A, B).cpu(), A @ B.T);"
        "A = torch.randn(257, 400, generator=g).double(); B = torch.randn(129, 400, generator=g).double();"
        "d = (_native.int8_matmul(A, B).cpu() - A @ B.T).abs().max().item(); assert d < 1e-12, d; print('ok')"
    ) % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "decoupled-kg_b200")
    env = dict(os.environ, DKG_OZ_PAIR="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 200, 416), (257, 129, 800)])
def test_nonnegative_b_takes_the_merged_n256_path_and_stays_exact(M, N, K):
    """With B >= 0 (as for kernel values) all B digit planes are unsigned and pairs of them are issued
    as single N = 256 MMAs; integers of up to 3 digits must still come out bit-exact."""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randint(-(2**20), 2**20, (M, K), generator=g).double()
    B = torch.randint(0, 2**20, (N, K), generator=g).double()
    want = A @ B.T  # |sum| < 2^50: exact in fp64
    assert torch.equal(_mm(A, B), want)
    Br = torch.rand(N, K, generator=g, dtype=torch.double)
    Ar = torch.randn(M, K, generator=g, dtype=torch.double)
    truth = (Ar.numpy().astype(np.longdouble) @ Br.numpy().astype(np.longdouble).T).astype(np.float64)
    err = np.abs(_mm(Ar, Br).numpy() - truth).max() / (np.abs(Ar.numpy()).max() * K)
    assert err < 4e-16, err
