"""CPU stand-in for ``dmd_era5_b200.device_ops.CudaOps`` - TEST INFRASTRUCTURE ONLY.

It lets the host-side drivers (rsvd.py, standard.py, dist.py) run on CPU tensors so that their
control flow, the delay-block bookkeeping and the gloo collectives can be tested without a GPU.
The product never imports this module; CudaOps refuses CPU tensors.
Semantics mirror include/era5svd.h; tall ops round to the tall dtype like the kernels do.
"""
from __future__ import annotations

import numpy as np
import torch


class FakeOps:
    name = "fake-cpu"

    def __init__(self):
        self.device = torch.device("cpu")
        self.calls: dict[str, int] = {}

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def to_device(self, host, non_blocking=True):
        return host.clone()

    def sketch(self, X, Om, Y=None, precision=0):
        self._count("sketch")
        out = (X.double() @ Om.double()).to(X.dtype)
        if Y is None:
            return out
        Y.copy_(out)
        return Y

    def project(self, X, Y, Z=None, accumulate=False, precision=0):
        self._count("project")
        out = X.double().t() @ Y.double()
        if Z is None:
            return out
        if accumulate:
            Z += out
        else:
            Z.copy_(out)
        return Z

    # -- tensor-core path stand-ins: exact float64 products of the float32 operands, rounded to float32 like the
    #    kernels' outputs; tf32 rounding emulated on the bit pattern (round to nearest, ties away: cvt.rna) --------
    @staticmethod
    def _tf32(a: torch.Tensor) -> torch.Tensor:
        bits = a.to(torch.float32).contiguous().view(torch.int32)
        return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)

    def tf32_ldy(self, l):
        return (l + 15) // 16 * 16

    def round_tf32_(self, A):
        self._count("round_tf32")
        A.copy_(self._tf32(A).double())
        return A

    def sketch_tf32x3(self, Xhi, Xlo, Om, Y, Yhi, Ylo, om_tf32=False):
        self._count("sketch_tc")
        X = Xhi.double() + (Xlo.double() if Xlo is not None else 0.0)
        om = self._tf32(Om).double() if om_tf32 else Om
        out = (X @ om).to(torch.float32)
        if Y is not None:
            Y.copy_(out)
        if Yhi is not None:
            hi = self._tf32(out)
            Yhi.copy_(hi)
            Ylo.copy_(out - hi)

    @staticmethod
    def _tf32_trunc(a: torch.Tensor) -> torch.Tensor:
        """what the tensor core does to a raw fp32 operand: the low 13 mantissa bits are dropped"""
        bits = a.to(torch.float32).contiguous().view(torch.int32)
        return (bits & ~0x1FFF).view(torch.float32)

    @staticmethod
    def _check_y(*ys):
        """the tensor-core projection kernels take their Y operand by TMA: 16-byte aligned base (include/era5svd.h)"""
        for y in ys:
            if y is not None and y.data_ptr() % 16:
                raise ValueError("project_tf32x3: Yhi / Ylo must be 16-byte aligned")

    def split_tf32(self, X):
        self._count("split_tf32")
        hi = self._tf32_trunc(X)
        return hi, (X - hi)

    def sketch_tf32x1(self, X, Om, Y):
        self._count("sketch_x1")
        Y.copy_((self._tf32_trunc(X).double() @ self._tf32(Om).double()).to(torch.float32))

    def project_tf32x1(self, X, Y, Z=None, accumulate=False):
        self._count("project_x1")
        self._check_y(Y)
        out = self._tf32_trunc(X).double().t() @ self._tf32_trunc(Y).double()
        if Z is None:
            return out
        if accumulate:
            Z += out
        else:
            Z.copy_(out)
        return Z

    def project_tf32x2(self, X, Y, Z=None, accumulate=False):
        self._count("project_x2")
        self._check_y(Y)
        out = X.double().t() @ self._tf32_trunc(Y).double()
        if Z is None:
            return out
        if accumulate:
            Z += out
        else:
            Z.copy_(out)
        return Z

    def project_tf32x3(self, Xhi, Xlo, Yhi, Ylo, Z=None, accumulate=False):
        self._count("project_tc")
        self._check_y(Yhi, Ylo)
        X = Xhi.double() + (Xlo.double() if Xlo is not None else 0.0)
        Yv = Yhi.double() + (Ylo.double() if Ylo is not None else 0.0)
        out = X.t() @ Yv
        if Z is None:
            return out
        if accumulate:
            Z += out
        else:
            Z.copy_(out)
        return Z

    def gemm(self, A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None):
        self._count("gemm")
        a = A.t() if transA else A
        b = B.t() if transB else B
        out = alpha * (a @ b)
        if C is None:
            return out.contiguous()
        C.copy_(out + beta * C)
        return C

    def syevj(self, A, max_sweeps=0, tol=0.0):
        self._count("syevj")
        w, v = np.linalg.eigh(0.5 * (A.numpy() + A.numpy().T))
        return torch.from_numpy(w[::-1].copy()), torch.from_numpy(np.ascontiguousarray(v[:, ::-1]))

    def chol_inv(self, G, rel_tol):
        """Right-looking Cholesky with the kernel's pivot rule (small_f64.cu chol_inv_kernel): a pivot that falls below
        rel_tol * G[j][j] (or is not positive) drops the direction - zero row of R, zero row / column of Rinv."""
        self._count("chol_inv")
        g = 0.5 * (G.numpy() + G.numpy().T)
        l = g.shape[0]
        A = g.copy()
        R = np.zeros_like(g)
        keep = []
        for j in range(l):
            d = A[j, j]
            if d > rel_tol * g[j, j] and d > 0.0:
                r = A[j, j:] / np.sqrt(d)
                R[j, j:] = r
                A[j:, j:] -= np.outer(r, r)
                keep.append(j)
        Rinv = np.zeros_like(g)
        if keep:
            ix = np.ix_(keep, keep)
            Rinv[ix] = np.linalg.inv(R[ix])
        return torch.from_numpy(R), torch.from_numpy(np.ascontiguousarray(Rinv))

    def col_normalize(self, P):
        self._count("col_normalize")
        nrm = torch.linalg.norm(P, dim=0)
        P /= torch.where(nrm > 0, nrm, torch.ones_like(nrm))       # a dropped (zero) column stays zero, like the kernel
        return nrm

    def sigma_from_eig(self, W):
        s = torch.sqrt(torch.clamp(W, min=0))
        inv = torch.where(s > 0, 1.0 / s, torch.zeros_like(s))
        return s, inv

    def convert(self, src, dtype):
        return src.to(dtype)

    def col_absmax(self, U, row_offset):
        self._count("col_absmax")
        a = U.double().abs()
        idx = torch.from_numpy(np.argmax(a.numpy(), axis=0))  # numpy argmax: FIRST maximum
        k = U.shape[1]
        vals = U.double()[idx, torch.arange(k)]
        return a[idx, torch.arange(k)], idx + row_offset, torch.sign(vals)

    def maxloc_combine(self, a, row, sgn):
        a, row, sgn = a.numpy(), row.numpy(), sgn.numpy()
        R, k = a.shape
        out = np.zeros(k)
        for c in range(k):
            best = max(range(R), key=lambda i: (a[i, c], -row[i, c]))
            out[c] = sgn[best, c]
        return torch.from_numpy(out)

    def scale_cols(self, U, scale):
        U *= scale.to(U.dtype)

    def scale_rows(self, V, scale):
        V *= scale[:, None]

    def build_rows(self, src, X, mean, std, weights, flags, nonfinite_flag=None):
        from oracle.slice_tools_np import standardize_np
        a = src.numpy()
        if flags & 1:
            out, mu, sd = standardize_np(a, scale=bool(flags & 2), axis=0)
            mean.copy_(torch.from_numpy(np.ascontiguousarray(mu)))
            if flags & 2:
                std.copy_(torch.from_numpy(np.ascontiguousarray(sd)))
        else:
            out = a
        out = torch.from_numpy(np.ascontiguousarray(out.T)).to(X.dtype)
        if weights is not None:
            out = out * weights[:, None]
        X.copy_(out)
