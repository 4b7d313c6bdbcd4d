"""Thin typed layer between torch tensor handles and the C ABI (``include/era5svd.h``).

torch is used only for device memory, streams and (in ``dist.py``) the process group: every
method below hands raw device pointers + leading dimensions to ``libera5svd.so`` on torch's
current CUDA stream.  Tensors must be CUDA, 2-D row-major views (``stride(1) == 1``); a
delay-embedded block is simply the view ``X[:, j : j + n]`` (same leading dimension).
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import F32, F64, PREC_NATIVE, PREC_TF32X3, check

_DT = {torch.float32: F32, torch.float64: F64}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; float32 / float64 only") from None


def _mat(t: torch.Tensor, name: str) -> tuple[int, int]:
    """pointer, leading dimension of a row-major CUDA matrix view."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (dmd_era5_b200 has no CPU path)")
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D, got shape {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError(f"{name} must be row-major (stride(1) == 1)")
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])
    return t.data_ptr(), ld


def _vec(t: torch.Tensor | None, name: str) -> int | None:
    if t is None:
        return None
    if not t.is_cuda or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous CUDA tensor")
    return t.data_ptr()


class KernelTimer:
    """CUDA-event timing of individual C-ABI calls on the launching stream (used by bench.py for the
    roofline numbers).  Events are only recorded; durations are read after a synchronize."""

    def __init__(self):
        self.records: list[tuple[str, torch.cuda.Event, torch.cuda.Event, dict]] = []

    def start(self, name: str, **meta):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        self.records.append((name, e0, e1, meta))
        return e1

    def summary(self) -> dict:
        out: dict[str, dict] = {}
        for name, e0, e1, meta in self.records:
            d = out.setdefault(name, {"calls": 0, "ms": 0.0, "bytes": 0.0, "flops": 0.0})
            d["calls"] += 1
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += meta.get("bytes", 0.0)
            d["flops"] += meta.get("flops", 0.0)
        return out


class CudaOps:
    """Kernel launcher bound to one device.  All calls are asynchronous on the current stream."""

    name = "cuda"

    def __init__(self, device: torch.device | str | int):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CudaOps needs a CUDA device")
        self.lib = _cabi.lib()
        self._index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", self._index)
        self._ws: dict[str, torch.Tensor] = {}
        self.timer: KernelTimer | None = None   # bench.py attaches one to time kernels with CUDA events
        self.reserve_small_pool()

    # -- memory ------------------------------------------------------------------------------
    def reserve_small_pool(self, nbytes: int = 32 << 20) -> None:
        """Pre-grow the caching allocator's small-block pool (2 MiB segments serving tensors <= 1 MiB: every small factor
        of the schedule).  Without it the pool grows lazily: with results of earlier calls still alive the 4th call of a
        series needs one more segment, and that cudaMalloc - issued while the GPU is busy - was measured to block the
        launching thread for 2 ... 88 ms, long enough to drain the launch queue (one 60 - 75 ms step in a run of 11 ms
        steps, profiles/r02_bench_stall.txt).  Call after torch.cuda.empty_cache(), which returns the reservation."""
        with torch.cuda.device(self.device):
            hold = [torch.empty(1 << 20, dtype=torch.uint8, device=self.device) for _ in range(max(2, int(nbytes) >> 20))]
            del hold

    def empty(self, shape, dtype) -> torch.Tensor:
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype) -> torch.Tensor:
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def to_device(self, host: torch.Tensor, non_blocking: bool = True) -> torch.Tensor:
        return host.to(self.device, non_blocking=non_blocking)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        buf = self._ws.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._ws[key] = buf
        return buf

    def _stream(self) -> int:
        # the C ABI launches on the calling thread's CURRENT device: a foreign current device would pair this
        # device's stream and pointers with another context - make it ours before handing the stream out
        if torch.cuda.current_device() != self._index:
            torch.cuda.set_device(self._index)
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- (a) matrix build --------------------------------------------------------------------
    def build_rows(self, src: torch.Tensor, X: torch.Tensor, mean: torch.Tensor | None,
                   std: torch.Tensor | None, weights: torch.Tensor | None, flags: int,
                   nonfinite_flag: torch.Tensor | None = None) -> None:
        """src: (T, P) native time-major view; X: (P, T) destination rows."""
        sp, sld = _mat(src, "src")
        xp, xld = _mat(X, "X")
        T, P = src.shape
        if tuple(X.shape) != (P, T):
            raise ValueError(f"X must be ({P}, {T}), got {tuple(X.shape)}")
        end = self.timer.start("build_rows", bytes=float(T * P) * (src.element_size() + X.element_size())) if self.timer else None
        check(self.lib.era5svd_build_rows(sp, _dt(src), T, sld, P, xp, _dt(X), xld, _vec(mean, "mean"),
                                          _vec(std, "std"), _vec(weights, "weights"), flags,
                                          _vec(nonfinite_flag, "flag"), self._stream()), "era5svd_build_rows")
        if end is not None:
            end.record()

    def build_rows_split(self, src: torch.Tensor, X: torch.Tensor | None, Xhi: torch.Tensor, Xlo: torch.Tensor,
                         mean: torch.Tensor | None, std: torch.Tensor | None, weights: torch.Tensor | None,
                         flags: int, nonfinite_flag: torch.Tensor | None = None) -> None:
        """Like build_rows for a float32 matrix, writing the tf32 hi / lo images in the same pass
        (X itself optional).  All outputs share one row pitch."""
        sp, sld = _mat(src, "src")
        T, P = src.shape
        hp, hld = _mat(Xhi, "Xhi"); lp, lld = _mat(Xlo, "Xlo")
        xp = None
        if X is not None:
            xp, xld = _mat(X, "X")
            if xld != hld:
                raise ValueError("build_rows_split: X, Xhi, Xlo must share their row pitch")
        if hld != lld or tuple(Xhi.shape) != (P, T) or Xhi.dtype != torch.float32:
            raise ValueError("build_rows_split: Xhi / Xlo must be float32 (P, T) with one row pitch")
        end = self.timer.start("build_rows", bytes=float(T * P) * (src.element_size() + 4)) if self.timer else None
        check(self.lib.era5svd_build_rows_split(sp, _dt(src), T, sld, P, xp, hp, lp, hld, _vec(mean, "mean"),
                                                _vec(std, "std"), _vec(weights, "weights"), flags,
                                                _vec(nonfinite_flag, "flag"), self._stream()), "era5svd_build_rows_split")
        if end is not None:
            end.record()

    def check_finite(self, X: torch.Tensor) -> torch.Tensor:
        """Device int32 flag (1,) = 1 if X holds a NaN / Inf (sklearn check_array, extmath.py:546); asynchronous."""
        flag = self.zeros((1,), torch.int32)
        xp, xld = _mat(X, "X")
        check(self.lib.era5svd_check_finite(xp, _dt(X), X.shape[0], X.shape[1], xld, flag.data_ptr(), self._stream()),
              "era5svd_check_finite")
        return flag

    # -- (b) tall passes ---------------------------------------------------------------------
    def sketch(self, X: torch.Tensor, Om: torch.Tensor, Y: torch.Tensor | None = None,
               precision: int = PREC_NATIVE) -> torch.Tensor:
        m, n = X.shape
        if Om.shape[0] != n or Om.dtype != X.dtype:
            raise ValueError("sketch: Om must be (n, l) with X's dtype")
        l = Om.shape[1]
        if Y is None:
            Y = self.empty((m, l), X.dtype)
        xp, xld = _mat(X, "X"); op, old = _mat(Om, "Om"); yp, yld = _mat(Y, "Y")
        es = X.element_size()
        end = self.timer.start("sketch" if n > 2 * l else "apply_basis", bytes=float(es) * (m * n + m * l + n * l),
                               flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_sketch(xp, _dt(X), m, n, xld, op, l, old, yp, yld, precision, self._stream()),
              "era5svd_sketch")
        if end is not None:
            end.record()
        return Y

    def project(self, X: torch.Tensor, Y: torch.Tensor, Z: torch.Tensor | None = None,
                accumulate: bool = False, precision: int = PREC_NATIVE) -> torch.Tensor:
        m, n = X.shape
        if Y.shape[0] != m or Y.dtype != X.dtype:
            raise ValueError("project: Y must be (m, l) with X's dtype")
        l = Y.shape[1]
        if Z is None:
            Z = self.empty((n, l), torch.float64)
            accumulate = False
        xp, xld = _mat(X, "X"); yp, yld = _mat(Y, "Y"); zp, zld = _mat(Z, "Z")
        nbytes = int(self.lib.era5svd_project_workspace_bytes(_dt(X), m, n, l, precision))
        ws = self._workspace("project", nbytes)
        es = X.element_size()
        end = self.timer.start("project" if n > 2 * l else "gram", bytes=float(es) * (m * n + m * l) + 8.0 * n * l,
                               flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_project(xp, _dt(X), m, n, xld, yp, l, yld, zp, zld, int(accumulate), precision,
                                       ws.data_ptr(), ws.numel(), self._stream()), "era5svd_project")
        if end is not None:
            end.record()
        return Z

    # -- (b') tensor-core passes (tcgen05, 3xTF32) ---------------------------------------------
    def split_tf32(self, X: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """hi / lo float32 images of X (same shape; row pitch padded to a multiple of 4 floats)."""
        if X.dtype != torch.float32:
            raise TypeError("split_tf32: float32 only")
        m, n = X.shape
        ld = -(-n // 4) * 4
        hi = self.empty((m, ld), torch.float32)[:, :n]
        lo = self.empty((m, ld), torch.float32)[:, :n]
        xp, xld = _mat(X, "X")
        end = self.timer.start("split_tf32", bytes=12.0 * m * n) if self.timer else None
        check(self.lib.era5svd_split_tf32(xp, m, n, xld, hi.data_ptr(), lo.data_ptr(), ld, self._stream()),
              "era5svd_split_tf32")
        if end is not None:
            end.record()
        return hi, lo

    @staticmethod
    def tf32_ldy(l: int) -> int:
        """Row pitch the tensor-core sketch needs for its outputs: round_up(l, 16)."""
        return -(-l // 16) * 16

    def round_tf32_(self, A: torch.Tensor) -> torch.Tensor:
        """A <- tf32(A) in place (float64 small factor): the values the tensor core reads exactly."""
        if A.dtype != torch.float64:
            raise ValueError("round_tf32_: float64 small factor expected")
        ap, ald = _mat(A, "A")
        check(self.lib.era5svd_round_tf32_f64(ap, A.shape[0], A.shape[1], ald, self._stream()), "era5svd_round_tf32_f64")
        return A

    def sketch_tf32x3(self, Xhi: torch.Tensor, Xlo: torch.Tensor, Om: torch.Tensor, Y: torch.Tensor | None,
                      Yhi: torch.Tensor | None, Ylo: torch.Tensor | None, om_tf32: bool = False) -> None:
        """Y (and/or the split pair Yhi, Ylo) = X @ Om; Om is the float64 (n, l) small factor.
        Outputs are (m, l) views of buffers whose row pitch is tf32_ldy(l).
        om_tf32: Om holds tf32-representable values (round_tf32_): two tensor-core products per k-step instead of
        three (era5svd_sketch_tf32x2; plain float32 X only)."""
        m, n = Xhi.shape
        l = Om.shape[1]
        if Om.dtype != torch.float64 or Om.shape[0] != n:
            raise ValueError("sketch_tf32x3: Om must be float64 (n, l)")
        # Xlo is None: Xhi is the plain float32 matrix and the tf32 split happens on chip (gemm_tc2.cu)
        hp, hld = _mat(Xhi, "Xhi"); op, old = _mat(Om, "Om")
        lp, lld = _mat(Xlo, "Xlo") if Xlo is not None else (None, hld)
        if (Yhi is None) != (Ylo is None):
            raise ValueError("sketch_tf32x3: Yhi and Ylo go together")
        outs = [t for t in (Y, Yhi, Ylo) if t is not None]
        lds = {_mat(t, "Y")[1] for t in outs}
        if hld != lld or len(lds) != 1:
            raise ValueError("sketch_tf32x3: hi/lo operands and outputs must share their row pitch")
        ldy = lds.pop()
        nbytes = int(self.lib.era5svd_sketch_tf32x3_workspace_bytes(n, l))
        ws = self._workspace("sketch_tc", nbytes)
        n_out = sum(t is not None for t in (Y, Yhi, Ylo))
        end = self.timer.start("sketch_tc" if n > 2 * l else "apply_basis_tc",
                               bytes=4.0 * (m * n * (2 if Xlo is not None else 1) + m * l * n_out + n * l),
                               flops=2.0 * m * n * l) if self.timer else None
        outp = (Y.data_ptr() if Y is not None else None, Yhi.data_ptr() if Yhi is not None else None,
                Ylo.data_ptr() if Ylo is not None else None)
        if om_tf32:
            if Xlo is not None:
                raise ValueError("sketch_tf32x3: om_tf32 needs the plain float32 matrix (Xlo=None)")
            check(self.lib.era5svd_sketch_tf32x2(hp, m, n, hld, op, l, old, *outp, ldy, ws.data_ptr(), ws.numel(),
                                                 self._stream()), "era5svd_sketch_tf32x2")
        else:
            check(self.lib.era5svd_sketch_tf32x3(hp, lp, m, n, hld, op, l, old, *outp, ldy, ws.data_ptr(), ws.numel(),
                                                 self._stream()), "era5svd_sketch_tf32x3")
        if end is not None:
            end.record()

    def sketch_tf32x1(self, X: torch.Tensor, Om: torch.Tensor, Y: torch.Tensor) -> None:
        """Y (plain float32, pitch tf32_ldy(l)) = X @ tf32(Om): ONE tensor-core product per k-step on the raw float32
        tiles (era5svd_sketch_tf32x1) - the early power iterations."""
        m, n = X.shape
        l = Om.shape[1]
        if Om.dtype != torch.float64 or Om.shape[0] != n:
            raise ValueError("sketch_tf32x1: Om must be float64 (n, l)")
        xp, xld = _mat(X, "X"); op, old = _mat(Om, "Om"); yp, yld = _mat(Y, "Y")
        ws = self._workspace("sketch_tc", int(self.lib.era5svd_sketch_tf32x3_workspace_bytes(n, l)))
        end = self.timer.start("sketch_x1", bytes=4.0 * (m * n + m * l + n * l), flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_sketch_tf32x1(xp, m, n, xld, op, l, old, yp, yld, ws.data_ptr(), ws.numel(), self._stream()),
              "era5svd_sketch_tf32x1")
        if end is not None:
            end.record()

    def project_tf32x2(self, X: torch.Tensor, Y: torch.Tensor, Z: torch.Tensor | None = None,
                       accumulate: bool = False) -> torch.Tensor:
        """Z (float64, n x l) (+)= X^T tf32(Y): X split on chip, Y truncated - two products per k-step
        (era5svd_project_tf32x2; the projection of the last power iteration)."""
        m, n = X.shape
        l = Y.shape[1]
        if Z is None:
            Z = self.empty((n, l), torch.float64)
            accumulate = False
        xp, xld = _mat(X, "X"); yp, yld = _mat(Y, "Y"); zp, zld = _mat(Z, "Z")
        ws = self._workspace("project", int(self.lib.era5svd_project_tf32x3_workspace_bytes(m, n, l)))
        end = self.timer.start("project_x2", bytes=4.0 * (m * n + m * l) + 8.0 * n * l, flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_project_tf32x2(xp, m, n, xld, yp, l, yld, zp, zld, int(accumulate), ws.data_ptr(),
                                              ws.numel(), self._stream()), "era5svd_project_tf32x2")
        if end is not None:
            end.record()
        return Z

    def project_tf32x1(self, X: torch.Tensor, Y: torch.Tensor, Z: torch.Tensor | None = None,
                       accumulate: bool = False) -> torch.Tensor:
        """Z (float64, n x l) (+)= X^T Y, one tensor-core product per k-step (era5svd_project_tf32x1)."""
        m, n = X.shape
        l = Y.shape[1]
        if Z is None:
            Z = self.empty((n, l), torch.float64)
            accumulate = False
        xp, xld = _mat(X, "X"); yp, yld = _mat(Y, "Y"); zp, zld = _mat(Z, "Z")
        ws = self._workspace("project", int(self.lib.era5svd_project_tf32x3_workspace_bytes(m, n, l)))
        end = self.timer.start("project_x1", bytes=4.0 * (m * n + m * l) + 8.0 * n * l, flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_project_tf32x1(xp, m, n, xld, yp, l, yld, zp, zld, int(accumulate), ws.data_ptr(),
                                              ws.numel(), self._stream()), "era5svd_project_tf32x1")
        if end is not None:
            end.record()
        return Z

    def project_tf32x3(self, Xhi: torch.Tensor, Xlo: torch.Tensor, Yhi: torch.Tensor, Ylo: torch.Tensor,
                       Z: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
        m, n = Xhi.shape
        l = Yhi.shape[1]
        if Z is None:
            Z = self.empty((n, l), torch.float64)
            accumulate = False
        hp, hld = _mat(Xhi, "Xhi")
        lp, lld = _mat(Xlo, "Xlo") if Xlo is not None else (None, hld)   # None: plain float32 X, split on chip
        yhp, yhld = _mat(Yhi, "Yhi"); zp, zld = _mat(Z, "Z")
        # Ylo None (with Xlo None): Yhi is the plain float32 Y of the sketch, split on chip as well
        ylp, ylld = _mat(Ylo, "Ylo") if Ylo is not None else (None, yhld)
        if Ylo is None and Xlo is not None:
            raise ValueError("project_tf32x3: a plain Y (Ylo=None) needs the plain X path (Xlo=None)")
        if hld != lld or yhld != ylld:
            raise ValueError("project_tf32x3: hi/lo operands must share their row pitch")
        nbytes = int(self.lib.era5svd_project_tf32x3_workspace_bytes(m, n, l))
        ws = self._workspace("project", nbytes)
        # algorithmic bytes: X once + the tall factor as it exists for this pass (ONE plain image in the power
        # iterations, the hi / lo PAIR in the final pass, where the Gram and U = Y M kernels consume the pair too)
        end = self.timer.start("project_tc" if n > 2 * l else "gram_tc",
                               bytes=4.0 * (m * n + m * l * (2 if Ylo is not None else 1)) + 8.0 * n * l,
                               flops=2.0 * m * n * l) if self.timer else None
        check(self.lib.era5svd_project_tf32x3(hp, lp, m, n, hld, yhp, ylp, l, yhld, zp, zld, int(accumulate),
                                              ws.data_ptr(), ws.numel(), self._stream()), "era5svd_project_tf32x3")
        if end is not None:
            end.record()
        return Z

    # -- (c)/(d) small float64 factors -------------------------------------------------------
    def gemm(self, A: torch.Tensor, B: torch.Tensor, transA: bool = False, transB: bool = False,
             alpha: float = 1.0, beta: float = 0.0, C: torch.Tensor | None = None) -> torch.Tensor:
        M = A.shape[1] if transA else A.shape[0]
        K = A.shape[0] if transA else A.shape[1]
        N = B.shape[0] if transB else B.shape[1]
        Kb = B.shape[1] if transB else B.shape[0]
        if K != Kb or A.dtype != torch.float64 or B.dtype != torch.float64:
            raise ValueError("gemm: shape / dtype mismatch (float64 only)")
        if C is None:
            C = self.empty((M, N), torch.float64)
            beta = 0.0
        ap, ald = _mat(A, "A"); bp, bld = _mat(B, "B"); cp, cld = _mat(C, "C")
        check(self.lib.era5svd_gemm_f64(int(transA), int(transB), M, N, K, alpha, ap, ald, bp, bld, beta, cp, cld,
                                        self._stream()), "era5svd_gemm_f64")
        return C

    def syevj(self, A: torch.Tensor, max_sweeps: int = 0, tol: float = 0.0) -> tuple[torch.Tensor, torch.Tensor]:
        """Eigen-decomposition of symmetric A (destroyed).  Returns (W descending, V columns)."""
        n = A.shape[0]
        ap, ald = _mat(A, "A")
        W = self.empty((n,), torch.float64)
        V = self.empty((n, n), torch.float64)
        nbytes = int(self.lib.era5svd_syevj_workspace_bytes(n))
        ws = self._workspace("syevj", nbytes)
        check(self.lib.era5svd_syevj_f64(ap, n, ald, W.data_ptr(), V.data_ptr(), n, max_sweeps, tol, ws.data_ptr(),
                                         ws.numel(), self._stream()), "era5svd_syevj_f64")
        return W, V

    # -- time-sized symmetric eigensolver (csrc/eig_tridiag.cu) -------------------------------
    def tridiag_reduce(self, A: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Householder tridiagonalisation of symmetric A in place (A then holds the reflectors).
        Returns (d (n,), e (n,), tau (n,))."""
        n = A.shape[0]
        ap, ald = _mat(A, "A")
        d = self.empty((n,), torch.float64); e = self.empty((n,), torch.float64); tau = self.empty((n,), torch.float64)
        ws = self._workspace("tridiag", int(self.lib.era5svd_tridiag_reduce_workspace_bytes(n)))
        end = self.timer.start("tridiag_reduce", bytes=8.0 * n ** 3, flops=4.0 / 3.0 * n ** 3) if self.timer else None
        check(self.lib.era5svd_tridiag_reduce_f64(ap, n, ald, d.data_ptr(), e.data_ptr(), tau.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), self._stream()), "era5svd_tridiag_reduce_f64")
        if end is not None:
            end.record()
        return d, e, tau

    def tridiag_eig_topk(self, d: torch.Tensor, e: torch.Tensor, k: int) -> tuple[torch.Tensor, torch.Tensor]:
        """k largest eigenvalues (descending) and (unorthogonalised) eigenvectors Z (n, k) of T = tridiag(e, d, e)."""
        n = d.shape[0]
        W = self.empty((k,), torch.float64)
        Z = self.empty((n, k), torch.float64)
        ws = self._workspace("tridiag_eig", int(self.lib.era5svd_tridiag_eig_topk_workspace_bytes(n, k)))
        check(self.lib.era5svd_tridiag_eig_topk_f64(d.data_ptr(), e.data_ptr(), n, k, W.data_ptr(), Z.data_ptr(), k,
                                                    ws.data_ptr(), ws.numel(), self._stream()),
              "era5svd_tridiag_eig_topk_f64")
        return W, Z

    def tridiag_apply(self, d: torch.Tensor, e: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
        n, k = Z.shape
        zp, zld = _mat(Z, "Z")
        Y = self.empty((n, k), torch.float64)
        check(self.lib.era5svd_tridiag_apply_f64(d.data_ptr(), e.data_ptr(), n, k, zp, zld, Y.data_ptr(), k,
                                                 self._stream()), "era5svd_tridiag_apply_f64")
        return Y

    def tridiag_backtransform(self, A: torch.Tensor, tau: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
        n, k = Z.shape
        ap, ald = _mat(A, "A"); zp, zld = _mat(Z, "Z")
        V = self.empty((n, k), torch.float64)
        check(self.lib.era5svd_tridiag_backtransform_f64(ap, n, ald, tau.data_ptr(), k, zp, zld, V.data_ptr(), k,
                                                         self._stream()), "era5svd_tridiag_backtransform_f64")
        return V

    def chol_inv(self, G: torch.Tensor, rel_tol: float) -> tuple[torch.Tensor, torch.Tensor]:
        l = G.shape[0]
        gp, gld = _mat(G, "G")
        R = self.empty((l, l), torch.float64)
        Rinv = self.empty((l, l), torch.float64)
        check(self.lib.era5svd_chol_inv_f64(gp, l, gld, R.data_ptr(), l, Rinv.data_ptr(), l, rel_tol, self._stream()),
              "era5svd_chol_inv_f64")
        return R, Rinv

    def col_normalize(self, P: torch.Tensor) -> torch.Tensor:
        n, l = P.shape
        pp, pld = _mat(P, "P")
        norms = self.empty((l,), torch.float64)
        check(self.lib.era5svd_col_normalize_f64(pp, n, l, pld, norms.data_ptr(), self._stream()),
              "era5svd_col_normalize_f64")
        return norms

    def sigma_from_eig(self, W: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        l = W.shape[0]
        s = self.empty((l,), torch.float64)
        inv = self.empty((l,), torch.float64)
        check(self.lib.era5svd_sigma_from_eig_f64(W.data_ptr(), l, s.data_ptr(), inv.data_ptr(), self._stream()),
              "era5svd_sigma_from_eig_f64")
        return s, inv

    def convert(self, src: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        if src.dtype == dtype:
            return src
        rows, cols = src.shape
        dst = self.empty((rows, cols), dtype)
        sp, sld = _mat(src, "src")
        check(self.lib.era5svd_convert(sp, _dt(src), sld, dst.data_ptr(), _DT[dtype], cols, rows, cols, self._stream()),
              "era5svd_convert")
        return dst

    # -- svd_flip ----------------------------------------------------------------------------
    def col_absmax(self, U: torch.Tensor, row_offset: int) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        m, k = U.shape
        up, uld = _mat(U, "U")
        a = self.empty((k,), torch.float64)
        row = self.empty((k,), torch.int64)
        sgn = self.empty((k,), torch.float64)
        nbytes = int(self.lib.era5svd_col_absmax_workspace_bytes(m, k))
        ws = self._workspace("absmax", nbytes)
        check(self.lib.era5svd_col_absmax(up, _dt(U), m, k, uld, row_offset, a.data_ptr(), row.data_ptr(),
                                          sgn.data_ptr(), ws.data_ptr(), ws.numel(), self._stream()),
              "era5svd_col_absmax")
        return a, row, sgn

    def maxloc_combine(self, a: torch.Tensor, row: torch.Tensor, sgn: torch.Tensor) -> torch.Tensor:
        """a, row, sgn: (R, k) stacked candidate sets -> sign (k,)."""
        R, k = a.shape
        a, row, sgn = a.contiguous(), row.contiguous(), sgn.contiguous()
        out = self.empty((k,), torch.float64)
        check(self.lib.era5svd_maxloc_combine(a.data_ptr(), row.data_ptr(), sgn.data_ptr(), R, k,
                                              out.data_ptr(), self._stream()), "era5svd_maxloc_combine")
        return out

    def scale_cols(self, U: torch.Tensor, scale: torch.Tensor) -> None:
        m, k = U.shape
        up, uld = _mat(U, "U")
        check(self.lib.era5svd_scale_cols(up, _dt(U), m, k, uld, scale.data_ptr(), self._stream()),
              "era5svd_scale_cols")

    def scale_rows(self, V: torch.Tensor, scale: torch.Tensor) -> None:
        k, n = V.shape
        vp, vld = _mat(V, "V")
        check(self.lib.era5svd_scale_rows_f64(vp, k, n, vld, scale.data_ptr(), self._stream()),
              "era5svd_scale_rows_f64")
