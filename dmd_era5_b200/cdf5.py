"""Classic-model NetCDF writer / reader in pure NumPy: CDF-2 (64-bit offsets) and CDF-5 (64-bit data).

Why it exists: the reference persists its result with ``to_netcdf(path, format="NETCDF4")``
(src/dmd_era5/era5_svd/era5_svd.py:434) through xarray + netCDF4, neither of which can be installed in this
image.  scipy's NetCDF-3 writer stops at 2 GiB per variable, which the headline configurations
exceed (c2 with ``save_data_matrix``: X = 3.1 GB; c3: U = 16 GB).  CDF-5 is the classic data model with every
size field widened to 64 bits; netCDF-C >= 4.4 (hence xarray's netcdf4 engine, ncdump, NCO, CDO) reads it as
``NETCDF3_64BIT_DATA``, so a file written here opens with ``xr.open_dataset`` on a machine that has the
reference's stack.  Layout written (NetCDF classic format specification):

    header   = magic numrecs dim_list gatt_list var_list         magic = 'C' 'D' 'F' \\x02 | \\x05
    dim      = name length                                        NON_NEG = int32 (CDF-2) | int64 (CDF-5)
    attr     = name nc_type nelems values(pad 4)
    var      = name ndims dimid* vatt_list nc_type vsize begin    begin: int64 in both versions
    data     = every (non-record) variable contiguous, big-endian, padded to 4 bytes

Only fixed-size variables are written (no record dimension: the stage knows every length up front).  Large arrays
are converted to big-endian chunk by chunk while they are written, so the peak extra memory is one chunk.
``version=2`` exists so that the header logic can be pinned byte for byte against scipy's writer
(tests/test_host_stage.py); the product writes version 5 whenever a variable does not fit NetCDF-3.
"""
from __future__ import annotations

import os
import struct

import numpy as np

NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
# nc_type codes: classic 1..6, CDF-5 adds 7..11
_TYPES = {"i1": 1, "S1": 2, "i2": 3, "i4": 4, "f4": 5, "f8": 6, "u1": 7, "u2": 8, "u4": 9, "i8": 10, "u8": 11}
_CODES = {v: k for k, v in _TYPES.items()}
CHUNK_BYTES = 256 << 20


def _nc_type(dt: np.dtype, version: int) -> int:
    key = "S1" if dt.kind == "S" else f"{dt.kind}{dt.itemsize}"
    code = _TYPES.get(key)
    if code is None or (code > 6 and version != 5):
        raise TypeError(f"dtype {dt} has no NetCDF classic (version {version}) external type")
    return code


class _Writer:
    def __init__(self, fp, version: int):
        if version not in (2, 5):
            raise ValueError("version must be 2 (64-bit offset) or 5 (64-bit data)")
        self.fp, self.version = fp, version
        self.nn = ">q" if version == 5 else ">i"        # NON_NEG

    def non_neg(self, v: int):
        self.fp.write(struct.pack(self.nn, int(v)))

    def tag(self, v: int):
        self.fp.write(struct.pack(">i", v))

    def name(self, s: str):
        b = s.encode("utf-8")
        self.non_neg(len(b))
        self.fp.write(b + b"\x00" * (-len(b) % 4))

    def att_values(self, v):
        if isinstance(v, (str, bytes)):
            b = v.encode("utf-8") if isinstance(v, str) else v
            self.tag(2)
            self.non_neg(len(b))
            self.fp.write(b + b"\x00" * (-len(b) % 4))
            return
        if isinstance(v, (bool, np.bool_)):
            v = int(v)
        arr = np.atleast_1d(np.asarray(v))
        if arr.dtype.kind in "iu" and not hasattr(v, "dtype"):       # Python ints: NC_INT when they fit
            arr = arr.astype(np.int32 if np.all(np.abs(arr) < 2 ** 31) else (np.int64 if self.version == 5 else np.float64))
        if arr.dtype == np.int64 and self.version != 5:
            arr = arr.astype(np.int32 if np.all(np.abs(arr) < 2 ** 31) else np.float64)
        if arr.dtype.kind == "f" and arr.dtype.itemsize not in (4, 8):
            arr = arr.astype(np.float64)
        code = _nc_type(arr.dtype, self.version)
        self.tag(code)
        self.non_neg(arr.size)
        b = arr.astype(arr.dtype.newbyteorder(">")).tobytes()
        self.fp.write(b + b"\x00" * (-len(b) % 4))

    def att_list(self, attrs: dict):
        if not attrs:
            self.tag(0)
            self.non_neg(0)
            return
        self.tag(NC_ATTRIBUTE)
        self.non_neg(len(attrs))
        for k, v in attrs.items():
            self.name(k)
            self.att_values(v)


def write_classic(path: str, dims: dict, variables: list, gattrs: dict, version: int = 5,
                  chunk_bytes: int = CHUNK_BYTES) -> None:
    """dims: name -> length (insertion order = dimension ids).
    variables: list of (name, dims tuple, array, attrs dict); arrays are numeric ndarrays (or np.memmap) or 'S1' arrays
    whose last dimension is a string-length dimension.  int64 data needs version 5."""
    dim_ids = {n: i for i, n in enumerate(dims)}
    prepared = []
    for name, vdims, arr, attrs in variables:
        arr = np.asarray(arr) if not isinstance(arr, np.memmap) else arr
        if tuple(arr.shape) != tuple(int(dims[d]) for d in vdims):
            raise ValueError(f"variable {name!r}: shape {tuple(arr.shape)} does not match its dimensions {vdims}")
        code = _nc_type(arr.dtype, version)
        nbytes = int(arr.size) * arr.dtype.itemsize
        vsize = nbytes + (-nbytes % 4)
        if version != 5 and vsize >= 2 ** 32 - 4:
            raise ValueError(f"variable {name!r} ({nbytes / 2 ** 30:.1f} GiB) needs the 64-bit-data format (version 5)")
        prepared.append((name, tuple(vdims), arr, dict(attrs or {}), code, nbytes, vsize))
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fp:
        w = _Writer(fp, version)
        fp.write(b"CDF" + bytes([version]))
        w.non_neg(0)                                         # numrecs: no record dimension
        if dims:
            w.tag(NC_DIMENSION)
            w.non_neg(len(dims))
            for n, length in dims.items():
                w.name(n)
                w.non_neg(length)
        else:
            w.tag(0); w.non_neg(0)
        w.att_list(gattrs)
        begin_pos = []
        if prepared:
            w.tag(NC_VARIABLE)
            w.non_neg(len(prepared))
            for name, vdims, arr, attrs, code, nbytes, vsize in prepared:
                w.name(name)
                w.non_neg(len(vdims))
                for d in vdims:
                    w.non_neg(dim_ids[d])
                w.att_list(attrs)
                w.tag(code)
                w.non_neg(vsize)
                begin_pos.append(fp.tell())
                fp.write(struct.pack(">q", 0))               # begin: patched below
        else:
            w.tag(0); w.non_neg(0)
        for (name, vdims, arr, attrs, code, nbytes, vsize), pos in zip(prepared, begin_pos):
            begin = fp.tell()
            fp.seek(pos)
            fp.write(struct.pack(">q", begin))
            fp.seek(begin)
            be = arr.dtype.newbyteorder(">") if arr.dtype.kind != "S" else arr.dtype
            if arr.ndim == 0 or nbytes <= chunk_bytes:
                fp.write(np.ascontiguousarray(arr, dtype=be).tobytes())
            else:
                # row chunks: ONE big-endian copy of a chunk at a time (the byte swap), written from that buffer directly
                # (no second ``tobytes`` copy); the swap of chunk i + 1 runs on a helper thread while chunk i is written
                # (NumPy releases the GIL in the conversion, the file write releases it too)
                from concurrent.futures import ThreadPoolExecutor

                rows = max(1, chunk_bytes // max(1, nbytes // arr.shape[0]))
                starts = list(range(0, arr.shape[0], rows))

                def swapped(r0: int) -> np.ndarray:
                    return np.ascontiguousarray(arr[r0 : r0 + rows], dtype=be).reshape(-1).view(np.uint8)

                with ThreadPoolExecutor(max_workers=1) as pool:
                    nxt = pool.submit(swapped, starts[0])
                    for i in range(len(starts)):
                        cur = nxt.result()
                        if i + 1 < len(starts):
                            nxt = pool.submit(swapped, starts[i + 1])
                        fp.write(cur.data)
            fp.write(b"\x00" * (vsize - nbytes))


class _Reader:
    def __init__(self, buf, version: int):
        self.buf, self.pos, self.version = buf, 4, version
        self.nn, self.nnsz = (">q", 8) if version == 5 else (">i", 4)

    def non_neg(self) -> int:
        v = struct.unpack_from(self.nn, self.buf, self.pos)[0]
        self.pos += self.nnsz
        return v

    def tag(self) -> int:
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def name(self) -> str:
        n = self.non_neg()
        s = bytes(self.buf[self.pos : self.pos + n]).decode("utf-8")
        self.pos += n + (-n % 4)
        return s

    def att_list(self) -> dict:
        tag, n = self.tag(), self.non_neg()
        out = {}
        if tag == 0:
            return out
        if tag != NC_ATTRIBUTE:
            raise ValueError("corrupt NetCDF header: attribute list expected")
        for _ in range(n):
            k = self.name()
            code, cnt = self.tag(), self.non_neg()
            dt = np.dtype(_CODES[code]).newbyteorder(">") if code != 2 else np.dtype("S1")
            nb = cnt * dt.itemsize
            raw = bytes(self.buf[self.pos : self.pos + nb])
            self.pos += nb + (-nb % 4)
            if code == 2:
                out[k] = raw.decode("utf-8")
            else:
                a = np.frombuffer(raw, dtype=dt).astype(dt.newbyteorder("="))
                out[k] = a[0].item() if a.size == 1 else a
        return out


def read_classic(path: str):
    """Header of a CDF-1/2/5 file + memory-mapped views of its fixed-size variables.
    Returns (dims: dict, variables: dict name -> (dims tuple, np.memmap in on-disk big-endian order, attrs), gattrs)."""
    with open(path, "rb") as fp:
        head = fp.read(4)
        if head[:3] != b"CDF" or head[3] not in (1, 2, 5):
            raise ValueError(f"{path}: not a NetCDF classic / 64-bit file")
        version = head[3]
    size = os.path.getsize(path)
    buf = np.memmap(path, dtype=np.uint8, mode="r", shape=(min(size, 64 << 20),))   # headers are small
    r = _Reader(buf, version)
    numrecs = r.non_neg()
    dims, order = {}, []
    tag, n = r.tag(), r.non_neg()
    if tag == NC_DIMENSION:
        for _ in range(n):
            nm = r.name()
            dims[nm] = r.non_neg()
            order.append(nm)
    gattrs = r.att_list()
    variables = {}
    tag, n = r.tag(), r.non_neg()
    specs = []
    if tag == NC_VARIABLE:
        for _ in range(n):
            nm = r.name()
            nd = r.non_neg()
            vdims = tuple(order[r.non_neg()] for _ in range(nd))
            attrs = r.att_list()
            code = r.tag()
            r.non_neg()                                      # vsize (recomputed from the shape)
            if version == 1:
                begin = struct.unpack_from(">i", buf, r.pos)[0]; r.pos += 4
            else:
                begin = struct.unpack_from(">q", buf, r.pos)[0]; r.pos += 8
            specs.append((nm, vdims, attrs, code, begin))
    del buf
    for nm, vdims, attrs, code, begin in specs:
        if vdims and dims[vdims[0]] == 0 and numrecs:
            raise NotImplementedError(f"{path}: record variable {nm!r} (this reader handles fixed-size variables)")
        dt = np.dtype(_CODES[code]).newbyteorder(">") if code != 2 else np.dtype("S1")
        shape = tuple(int(dims[d]) for d in vdims)
        arr = np.memmap(path, dtype=dt, mode="r", offset=begin, shape=shape) if int(np.prod(shape, dtype=np.int64)) > 0 \
            else np.zeros(shape, dtype=dt)
        variables[nm] = (vdims, arr, attrs)
    return dims, variables, gattrs
