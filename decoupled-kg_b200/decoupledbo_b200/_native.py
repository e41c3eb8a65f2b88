"""ctypes binding of ``libdkg_b200.so`` (C-ABI declared in ``include/dkg_b200.h``).

There is NO fallback: if the library is missing or no CUDA device is visible, the calls raise.
PyTorch is used only for device memory, streams and (elsewhere) ``torch.distributed``.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_int32, c_int64, c_uint32, c_void_p
from typing import Optional, Sequence

import torch
from torch import Tensor

from .gp_state import GPModelList

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "lib", "libdkg_b200.so")

DKG_OK, DKG_EINVAL, DKG_ECUDA, DKG_ENOTPD, DKG_ENOMEM, DKG_EEMPTY, DKG_ETRUNC, DKG_ECAPACITY = 0, -1, -2, -3, -4, -5, -6, -7
ABI_VERSION = 4

EXPORTED_SYMBOLS = (
    "dkg_abi_version",
    "dkg_last_error",
    "dkg_plan_create",
    "dkg_plan_destroy",
    "dkg_plan_append_point",
    "dkg_forward_dev",
    "dkg_forward_host",
    "dkg_expected_max_lines_dev",
    "dkg_piecewise_expectation_dev",
    "dkg_posterior_mean_dev",
    "dkg_int8_matmul_dev",
    "dkg_int8_peak",
    "dkg_plan_read",
    "dkg_launch_count",
    "dkg_launch_count_reset",
    "dkg_plan_stats",
    "dkg_profile_enable",
    "dkg_profile_read",
)
PROFILE_CATEGORIES = (
    "xprep", "gemm_T", "var", "gemm_cov", "place_own", "zstat", "filter", "hull", "overflow", "finalize", "digits",
)


class NativeLibraryError(RuntimeError):
    pass


class PlanCapacityError(RuntimeError):
    """``Plan.append_point``: no room left for another training point -- build a new plan."""


class _Objective(Structure):
    _fields_ = [
        ("train_x_dev", c_void_p),
        ("train_y_dev", c_void_p),
        ("n", c_int32),
        ("kernel", c_int32),
        ("lengthscale_host", POINTER(c_double)),
        ("outputscale", c_double),
        ("mean_const", c_double),
        ("noise", c_double),
        ("y_mean", c_double),
        ("y_std", c_double),
    ]


_lib = None


def load_library() -> ctypes.CDLL:
    """Load ``libdkg_b200.so`` (built in-tree by ``decoupled-kg_b200/build.sh``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: build it with `decoupled-kg_b200/build.sh` "
            f"(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            f"There is no CPU fallback for the discrete-KG hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    lib.dkg_abi_version.restype = ctypes.c_int
    lib.dkg_last_error.restype = c_char_p
    lib.dkg_plan_create.restype = ctypes.c_int
    lib.dkg_plan_create.argtypes = [
        POINTER(_Objective), c_int32, c_int32, c_void_p, c_int32, POINTER(c_double), c_int32,
        c_int32, c_uint32, c_void_p, POINTER(c_void_p),
    ]
    lib.dkg_plan_destroy.restype = None
    lib.dkg_plan_destroy.argtypes = [c_void_p]
    lib.dkg_plan_append_point.restype = ctypes.c_int
    lib.dkg_plan_append_point.argtypes = [c_void_p, c_int32, POINTER(c_double), c_double, c_void_p]
    lib.dkg_forward_dev.restype = ctypes.c_int
    lib.dkg_forward_dev.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.dkg_forward_host.restype = ctypes.c_int
    lib.dkg_forward_host.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.dkg_expected_max_lines_dev.restype = ctypes.c_int
    lib.dkg_expected_max_lines_dev.argtypes = [
        c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
        c_void_p, c_void_p, c_void_p,
    ]
    lib.dkg_piecewise_expectation_dev.restype = ctypes.c_int
    lib.dkg_piecewise_expectation_dev.argtypes = [
        c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
    ]
    lib.dkg_posterior_mean_dev.restype = ctypes.c_int
    lib.dkg_posterior_mean_dev.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
    lib.dkg_int8_matmul_dev.restype = ctypes.c_int
    lib.dkg_int8_matmul_dev.argtypes = [
        c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
        c_int32, c_void_p,
    ]
    lib.dkg_int8_peak.restype = ctypes.c_int
    lib.dkg_int8_peak.argtypes = [c_int32, c_int32, c_int32, c_int32, c_int32, POINTER(c_double), POINTER(c_double), c_void_p]
    lib.dkg_plan_read.restype = c_int64
    lib.dkg_plan_read.argtypes = [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]
    lib.dkg_launch_count.restype = c_int64
    lib.dkg_launch_count_reset.restype = None
    lib.dkg_plan_stats.restype = ctypes.c_int
    lib.dkg_plan_stats.argtypes = [c_void_p, POINTER(c_int64), c_void_p]
    lib.dkg_profile_enable.restype = None
    lib.dkg_profile_enable.argtypes = [ctypes.c_int]
    lib.dkg_profile_read.restype = ctypes.c_int
    lib.dkg_profile_read.argtypes = [POINTER(c_double), POINTER(c_int64), c_int32]
    if lib.dkg_abi_version() != ABI_VERSION:
        raise NativeLibraryError(
            f"ABI mismatch: library {lib.dkg_abi_version()} vs binding {ABI_VERSION}; rebuild."
        )
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc == DKG_OK:
        return
    msg = load_library().dkg_last_error().decode("utf-8", "replace")
    if rc == DKG_EEMPTY:
        raise ValueError(msg)
    if rc == DKG_EINVAL:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed ({rc}): {msg}")


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise NativeLibraryError(
            "no CUDA device visible: the discrete-KG hot path runs only on the GPU "
            "(libdkg_b200, sm_100a); there is no CPU fallback."
        )
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[Tensor]) -> c_void_p:
    return c_void_p(0 if t is None else t.data_ptr())


def launch_count() -> int:
    return int(load_library().dkg_launch_count())


def launch_count_reset() -> None:
    load_library().dkg_launch_count_reset()


def profile_enable(on: bool) -> None:
    load_library().dkg_profile_enable(1 if on else 0)


def profile_read():
    """-> {category: (total_ms, launches)} since the last read (synchronises the device)."""
    n = len(PROFILE_CATEGORIES)
    ms = (c_double * n)()
    cnt = (c_int64 * n)()
    _check(load_library().dkg_profile_read(ms, cnt, n), "dkg_profile_read")
    return {name: (float(ms[k]), int(cnt[k])) for k, name in enumerate(PROFILE_CATEGORIES)}


PLAN_DEFAULT = 0
PLAN_FAST32 = 1  # DKG_PLAN_FAST32: 4-digit covariance contraction (float32-class accuracy of that term)


class Plan:
    """Candidate-independent state of one acquisition function (``dkg_plan``)."""

    def __init__(
        self,
        model: GPModelList,
        x_discretisation: Tensor,
        scalarisation_weights: Tensor,
        target_output_ix: Optional[int],
        flags: int = 0,
    ):
        self._handle = c_void_p(0)
        lib = load_library()
        dev = require_cuda()
        self.device = dev
        M = model.num_outputs
        d = model.models[0].d
        xd = x_discretisation.detach().to(device=dev, dtype=torch.double).contiguous()
        if xd.dim() != 2 or xd.shape[1] != d:
            raise ValueError(f"x_discretisation must be (N, {d}); got {tuple(xd.shape)}")
        W = scalarisation_weights.detach().to(device="cpu", dtype=torch.double).contiguous()
        if W.dim() != 2 or W.shape[1] != M:
            raise ValueError(f"scalarisation_weights must be (S, {M}); got {tuple(W.shape)}")
        self.N, self.d, self.M, self.S = xd.shape[0], d, M, W.shape[0]
        self.target = -1 if target_output_ix is None else int(target_output_ix)  # -1 = coupled
        objs = (_Objective * M)()
        keep = []
        for m, o in enumerate(model.models):
            tx = o.train_x.to(device=dev, dtype=torch.double).contiguous()
            ty = o.train_y.to(device=dev, dtype=torch.double).contiguous()
            ls = (c_double * d)(*[float(v) for v in o.lengthscale.tolist()])
            keep += [tx, ty, ls]
            objs[m].train_x_dev = tx.data_ptr()
            objs[m].train_y_dev = ty.data_ptr()
            objs[m].n = o.n
            objs[m].kernel = o.kernel
            objs[m].lengthscale_host = ctypes.cast(ls, POINTER(c_double))
            objs[m].outputscale = o.outputscale
            objs[m].mean_const = o.mean_const
            objs[m].noise = o.noise
            objs[m].y_mean = o.y_mean
            objs[m].y_std = o.y_std
        self.n_target = model.models[self.target].n if 0 <= self.target < M else 0
        self.n_total = sum(o.n for o in model.models)
        w_host = (c_double * (W.numel()))(*W.reshape(-1).tolist())
        handle = c_void_p(0)
        with torch.cuda.device(dev):
            rc = lib.dkg_plan_create(
                objs, M, d, _ptr(xd), self.N, ctypes.cast(w_host, POINTER(c_double)), self.S,
                self.target, int(flags), _stream_ptr(), byref(handle),
            )
        _check(rc, "dkg_plan_create")
        self._handle = handle
        self._keep = keep  # the library copies what it needs, but keep inputs alive until here

    def append_point(self, m: int, x: Tensor, y: float) -> None:
        """Incremental refresh (``dkg_plan_append_point``): objective ``m`` received the observation
        ``(x, y)`` (``y`` in the space of ``train_y``) and the hyper-parameters are unchanged -- what the
        reference's BO loop does between iterations (``bo_loop.py:403-405``, fixed hyper-parameters
        ``:574-589``).  O(n^2 + n N) instead of a new plan's O(n^3 + n^2 N).  Raises
        ``PlanCapacityError`` when the plan has no room left (then build a new one)."""
        xs = [float(v) for v in torch.as_tensor(x, dtype=torch.double).reshape(-1).tolist()]
        if len(xs) != self.d:
            raise ValueError(f"x must have {self.d} coordinates; got {len(xs)}")
        xb = (c_double * self.d)(*xs)
        with torch.cuda.device(self.device):
            rc = load_library().dkg_plan_append_point(self._handle, int(m), xb, float(y), _stream_ptr())
        if rc == DKG_ECAPACITY:
            raise PlanCapacityError(load_library().dkg_last_error().decode("utf-8", "replace"))
        _check(rc, "dkg_plan_append_point")
        self.n_total += 1
        if int(m) == self.target:
            self.n_target += 1

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            load_library().dkg_plan_destroy(self._handle)
            self._handle = c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- forward ---------------------------------------------------------------------------
    def forward_device(self, X: Tensor, need_grad: bool, out_kg: Optional[Tensor] = None,
                       out_dX: Optional[Tensor] = None):
        """X: (C, d) float64 CUDA tensor -> (kg (C,), dX (C, d) | None) on the same device.
        ``out_kg`` / ``out_dX`` (contiguous CUDA float64 tensors of those shapes) receive the results
        in place, e.g. slices of a packed buffer that is all-gathered afterwards."""
        assert X.is_cuda and X.dtype == torch.double and X.dim() == 2 and X.shape[1] == self.d
        X = X.contiguous()
        C = X.shape[0]
        if out_kg is not None:
            assert out_kg.is_cuda and out_kg.dtype == torch.double and out_kg.is_contiguous() and out_kg.numel() == C
        if out_dX is not None:
            assert out_dX.is_cuda and out_dX.dtype == torch.double and out_dX.is_contiguous() and out_dX.numel() == C * self.d
        kg = out_kg if out_kg is not None else torch.empty(C, dtype=torch.double, device=X.device)
        dX = None
        if need_grad:
            dX = out_dX if out_dX is not None else torch.empty(C, self.d, dtype=torch.double, device=X.device)
        with torch.cuda.device(X.device):
            rc = load_library().dkg_forward_dev(self._handle, _ptr(X), C, _ptr(kg), _ptr(dX), _stream_ptr())
        _check(rc, "dkg_forward_dev")
        return kg, dX

    def forward_host(self, X: Tensor, need_grad: bool, out_kg: Optional[Tensor] = None,
                     out_dX: Optional[Tensor] = None):
        """X: (C, d) float64 CPU tensor -> (kg, dX | None) CPU tensors; copies inside the call."""
        assert (not X.is_cuda) and X.dtype == torch.double and X.dim() == 2 and X.shape[1] == self.d
        X = X.contiguous()
        C = X.shape[0]
        if out_kg is None and out_dX is None and C * (1 + self.d) <= 8192:
            # small batch (the optimiser's restarts): the few hundred bytes go straight into fresh
            # tensors -- a pinned staging copy plus a clone costs more than the transfer
            kg = torch.empty(C, dtype=torch.double)
            dX = torch.empty(C, self.d, dtype=torch.double) if need_grad else None
            with torch.cuda.device(self.device):
                rc = load_library().dkg_forward_host(self._handle, _ptr(X), C, _ptr(kg), _ptr(dX), _stream_ptr())
            _check(rc, "dkg_forward_host")
            return kg, dX
        # results land in pinned staging buffers owned by the plan (fast D2H), then are copied
        # into fresh tensors so callers may keep them across calls
        if getattr(self, "_pin_cap", 0) < C:
            self._pin_kg = torch.empty(C, dtype=torch.double).pin_memory()
            self._pin_dX = torch.empty(C, self.d, dtype=torch.double).pin_memory()
            self._pin_cap = C
        kg_pin = self._pin_kg[:C]
        dX_pin = self._pin_dX[:C] if need_grad else None
        with torch.cuda.device(self.device):
            rc = load_library().dkg_forward_host(self._handle, _ptr(X), C, _ptr(kg_pin), _ptr(dX_pin), _stream_ptr())
        _check(rc, "dkg_forward_host")
        kg = kg_pin.clone() if out_kg is None else out_kg.copy_(kg_pin)
        dX = None
        if need_grad:
            dX = dX_pin.clone() if out_dX is None else out_dX.copy_(dX_pin)
        return kg, dX

    def posterior_mean(self, X: Tensor) -> Tensor:
        """Posterior means of all objectives at the rows of X (C, d) -> (C, M) on X's device
        (``BoTorchModel.batch_fitness`` of the reference's metrics, pareto/sample.py:138-144)."""
        on_host = not X.is_cuda
        Xd = X.detach().to(device=self.device, dtype=torch.double).contiguous()
        mu = torch.empty(Xd.shape[0], self.M, dtype=torch.double, device=self.device)
        with torch.cuda.device(self.device):
            rc = load_library().dkg_posterior_mean_dev(self._handle, _ptr(Xd), Xd.shape[0], _ptr(mu), _stream_ptr())
        _check(rc, "dkg_posterior_mean_dev")
        return mu.cpu() if on_host else mu

    # -- introspection ---------------------------------------------------------------------
    _SHAPES = {
        "B": lambda s, C: (s.n_target, s.N),
        "Kinv": lambda s, C: (s.n_target, s.n_target),
        "chol": lambda s, C: (s.n_target, s.n_target),
        "alpha": lambda s, C: (s.n_total,),
        "mu_disc": lambda s, C: (s.N, s.M),
        "A0": lambda s, C: (s.S, s.N),
        "A0max": lambda s, C: (s.S,),
        "slopes": lambda s, C: (C, s.N + 1),
        "a_new": lambda s, C: (C, s.S),
        "var": lambda s, C: (C,),
        "kg_terms": lambda s, C: (C, s.S),
    }

    def read(self, name: str) -> Tensor:
        lib = load_library()
        with torch.cuda.device(self.device):
            count = lib.dkg_plan_read(self._handle, name.encode(), c_void_p(0), 0, _stream_ptr())
            if count < 0:
                _check(int(count), f"dkg_plan_read({name})")
            out = torch.empty(int(count), dtype=torch.double, device=self.device)
            rc = lib.dkg_plan_read(self._handle, name.encode(), _ptr(out), count, _stream_ptr())
            if rc < 0:
                _check(int(rc), f"dkg_plan_read({name})")
        stats = self.stats()
        shape = self._SHAPES[name](self, stats[0])
        return out.reshape(shape)

    def stats(self):
        buf = (c_int64 * 8)()
        with torch.cuda.device(self.device):
            rc = load_library().dkg_plan_stats(self._handle, buf, _stream_ptr())
        _check(rc, "dkg_plan_stats")
        return [int(v) for v in buf]


def int8_matmul(A: Tensor, Bt: Tensor, n_digits: int = 0, n_diagonals: int = 0) -> Tensor:
    """``A @ Bt.T`` in float64 accuracy on the int8 tensor cores (``dkg_int8_matmul_dev``): the
    contraction engine of the covariance rows, exposed for its own parity tests."""
    dev = require_cuda()
    A = A.detach().to(device=dev, dtype=torch.double).contiguous()
    Bt = Bt.detach().to(device=dev, dtype=torch.double).contiguous()
    if A.dim() != 2 or Bt.dim() != 2 or A.shape[1] != Bt.shape[1]:
        raise ValueError(f"A (M, K) and Bt (N, K) expected; got {tuple(A.shape)}, {tuple(Bt.shape)}")
    M, K = A.shape
    N = Bt.shape[0]
    D = torch.empty(M, N, dtype=torch.double, device=dev)
    with torch.cuda.device(dev):
        rc = load_library().dkg_int8_matmul_dev(
            _ptr(A), K, _ptr(Bt), K, M, N, K, n_digits, n_diagonals, _ptr(D), N, _stream_ptr()
        )
    _check(rc, "dkg_int8_matmul_dev")
    return D


def expected_max_lines(a: Tensor, b: Tensor, hull_cap: int = 64, want_grad: bool = False):
    """Device equivalent of ``calculate_epigraph_indices`` + ``calculate_expected_value_of_
    piecewise_linear_function`` (discretekg.py:341-452) for P line sets: a, b are (P, L) float64.

    Returns dict(emax (P,), hull_count (P,), hull_idx (P, cap), hull_x (P, cap), [dE_da, dE_db (P, L)]).
    """
    dev = require_cuda()
    a = a.detach().to(device=dev, dtype=torch.double).contiguous()
    b = b.detach().to(device=dev, dtype=torch.double).contiguous()
    if a.dim() != 2 or a.shape != b.shape:
        raise ValueError(f"a and b must be (P, L) with equal shapes; got {tuple(a.shape)}, {tuple(b.shape)}")
    P, L = a.shape
    emax = torch.empty(P, dtype=torch.double, device=dev)
    cnt = torch.zeros(P, dtype=torch.int32, device=dev)
    idx = torch.full((P, max(hull_cap, 1)), -1, dtype=torch.int32, device=dev)
    hx = torch.full((P, max(hull_cap, 1)), float("nan"), dtype=torch.double, device=dev)
    da = torch.empty(P, L, dtype=torch.double, device=dev) if want_grad else None
    db = torch.empty(P, L, dtype=torch.double, device=dev) if want_grad else None
    with torch.cuda.device(dev):
        rc = load_library().dkg_expected_max_lines_dev(
            _ptr(a), _ptr(b), P, L, _ptr(emax), _ptr(cnt), _ptr(idx), _ptr(hx), hull_cap,
            _ptr(da), _ptr(db), _stream_ptr(),
        )
    _check(rc, "dkg_expected_max_lines_dev")
    out = dict(emax=emax, hull_count=cnt, hull_idx=idx, hull_x=hx)
    if want_grad:
        out["dE_da"], out["dE_db"] = da, db
    return out


def piecewise_expectation(a: Tensor, b: Tensor, z: Tensor, want_grad: bool = False):
    """Device equivalent of ``calculate_expected_value_of_piecewise_linear_function``
    (discretekg.py:415-452) for P functions: a, b are (P, H), the break points z are (P, H-1).

    Returns dict(e (P,), [dE_da (P, H), dE_db (P, H), dE_dz (P, H-1)]) on the CUDA device.
    """
    dev = require_cuda()
    a = a.detach().to(device=dev, dtype=torch.double).contiguous()
    b = b.detach().to(device=dev, dtype=torch.double).contiguous()
    z = z.detach().to(device=dev, dtype=torch.double).contiguous()
    if a.dim() != 2 or a.shape != b.shape or z.shape != (a.shape[0], max(a.shape[1] - 1, 0)):
        raise ValueError(
            f"a, b must be (P, H) and z (P, H-1); got {tuple(a.shape)}, {tuple(b.shape)}, {tuple(z.shape)}"
        )
    P, H = a.shape
    e = torch.empty(P, dtype=torch.double, device=dev)
    da = torch.empty(P, H, dtype=torch.double, device=dev) if want_grad else None
    db = torch.empty(P, H, dtype=torch.double, device=dev) if want_grad else None
    dz = torch.empty(P, max(H - 1, 0), dtype=torch.double, device=dev) if want_grad else None
    with torch.cuda.device(dev):
        rc = load_library().dkg_piecewise_expectation_dev(
            _ptr(a), _ptr(b), _ptr(z), P, H, _ptr(e), _ptr(da), _ptr(db), _ptr(dz), _stream_ptr()
        )
    _check(rc, "dkg_piecewise_expectation_dev")
    out = dict(e=e)
    if want_grad:
        out["dE_da"], out["dE_db"], out["dE_dz"] = da, db, dz
    return out


def int8_peak(M: int, N: int, K: int, reps: int = 10, mode: int = 3):
    """Measured int8 tensor peak of the library's own MMA stream (``dkg_int8_peak``): (TOP/s, ms per launch).
    mode 3 = MMA stream only; 1 = no operand copies; 2 = no accumulator drains; 0 = full kernel on dummy digits."""
    dev = require_cuda()
    tops, ms = c_double(0.0), c_double(0.0)
    with torch.cuda.device(dev):
        rc = load_library().dkg_int8_peak(M, N, K, reps, mode, byref(tops), byref(ms), _stream_ptr())
    _check(rc, "dkg_int8_peak")
    return float(tops.value), float(ms.value)
