"""Minimal reader for R's .RData (RDX3, XDR) files holding the reference's bundled `dgCMatrix` fixtures
(`data/intestinalData.RData`, `data/intestinalDataSmall.RData`; docs R/intestinalData.R:1-11,
R/intestinalDataSmall.R:12-23).  Host-side plumbing only: it turns the fixture into numpy arrays so the
hot path can be exercised on the reference's own data without an R interpreter.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np


class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.p = 0
        self.refs: list = []

    def i32(self) -> int:
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n: int) -> bytes:
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == 254:  # NILVALUE_SXP
            return None
        if t in (253, 252, 251, 250, 249, 248, 247, 242, 241):  # env / missing / unbound singletons
            return ("special", t)
        if t == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (2, 6, 239, 240):  # LISTSXP / LANGSXP (+ATTR variants)
            out = []
            cur_flags, cur_t = flags, t
            while True:
                attr = self.item() if (cur_flags & (1 << 9)) else None
                tag = self.item() if (cur_flags & (1 << 10)) else None
                car = self.item()
                out.append((tag, car))
                nxt = self.i32()
                nt = nxt & 0xFF
                if nt == 254:
                    break
                if nt not in (2, 6, 239, 240):
                    self.p -= 4
                    out.append(("__cdr__", self.item()))
                    break
                cur_flags, cur_t = nxt, nt
            return out
        if t == 9:  # CHARSXP
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.p).astype(np.int32)
            self.p += 4 * n
            return self._with_attr(v, has_attr)
        if t == 14:  # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.p).astype(np.float64)
            self.p += 8 * n
            return self._with_attr(v, has_attr)
        if t == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        if t in (19, 20):  # VECSXP / EXPRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
            return self._with_attr(v, has_attr)
        if t == 25:  # S4SXP
            attrs = self.item() if has_attr else []
            return {"__s4__": True, **{k: v for k, v in attrs if k is not None}}
        if t == 238:  # ALTREP_SXP
            info = self.item()
            state = self.item()
            attr = self.item()
            return ("altrep", info, state, attr)
        raise NotImplementedError(f"SEXP type {t} at offset {self.p}")

    def _with_attr(self, v, has_attr):
        if has_attr:
            attrs = self.item()
            return {"__value__": v, "__attr__": {k: a for k, a in attrs if k is not None}}
        return v


def load_rdata(path: str) -> dict:
    """Returns {object name: parsed object}.  S4 objects come back as dicts of slots."""
    raw = open(path, "rb").read()
    for dec in (bz2.decompress, gzip.decompress, lzma.decompress, lambda b: b):
        try:
            buf = dec(raw)
            if buf[:4] in (b"RDX3", b"RDX2"):
                break
        except Exception:
            continue
    else:
        raise ValueError(f"{path}: not an RDX2/RDX3 file")
    r = _Reader(buf)
    r.p = 5  # "RDX3\n"
    if r.raw(2) != b"X\n":
        raise ValueError("only XDR serialisation is supported")
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.raw(n)
    top = r.item()
    return {k: v for k, v in top if k is not None}


def _unwrap(v):
    return v["__value__"] if isinstance(v, dict) and "__value__" in v else v


def load_dgcmatrix(path: str, name: str | None = None):
    """Returns (dense float64 matrix genes x cells, gene names, cell names)."""
    objs = load_rdata(path)
    obj = objs[name] if name else next(iter(objs.values()))
    i = _unwrap(obj["i"])
    p = _unwrap(obj["p"])
    x = _unwrap(obj["x"])
    dim = _unwrap(obj["Dim"])
    dn = _unwrap(obj["Dimnames"])
    nr, nc = int(dim[0]), int(dim[1])
    M = np.zeros((nr, nc), dtype=np.float64)
    cols = np.repeat(np.arange(nc), np.diff(p))
    M[i, cols] = x
    genes = list(_unwrap(dn[0])) if dn[0] is not None else [f"g{j}" for j in range(nr)]
    cells = list(_unwrap(dn[1])) if dn[1] is not None else [f"c{j}" for j in range(nc)]
    return M, genes, cells
