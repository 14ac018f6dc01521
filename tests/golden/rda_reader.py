"""Minimal reader for R `save()` files (RDX2 / XDR serialization, version 2).

Test infrastructure only: used by `make_fixtures.py` (run once, in the build
container where /root/reference exists) to convert the reference's bundled
datasets (`/root/reference/data/*.rda`, documented in `/root/reference/R/data.R:1-53`)
into the `.npz` fixtures committed next to this file.  The GPU box has no
/root/reference, so tests read only the `.npz` files.

Format notes (R internals manual, "Serialization formats"): optional gzip/bzip2/xz
wrapper, magic `RDX2\\n`, format byte `X\\n` (XDR = big-endian), three ints
(version, writer, min reader), then one serialized pairlist whose tags are the
saved object names.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE_SXP, REFSXP = 254, 255
SYMSXP, LISTSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, STRSXP, VECSXP = 1, 2, 9, 10, 13, 14, 16, 19
NA_INT = -2147483648


class RObj:
    """A vector-ish R object: `.value` (numpy array / list) plus `.attr` dict."""

    def __init__(self, value, attr=None):
        self.value = value
        self.attr = attr or {}

    def __repr__(self):
        return f"RObj({type(self.value).__name__}, attr={list(self.attr)})"


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        has_tag = bool(flags & (1 << 10))
        if t == NILVALUE_SXP:
            return None
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == SYMSXP:
            name = self.item()
            self.refs.append(name)
            return name
        if t == LISTSXP:
            out = []
            while True:
                attr = self.item() if has_attr else None  # noqa: F841 (pairlist attrs unused)
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                has_attr = bool(flags & (1 << 9))
                has_tag = bool(flags & (1 << 10))
                if t == NILVALUE_SXP:
                    break
                if t != LISTSXP:
                    raise ValueError(f"pairlist cdr of type {t}")
            return out
        if t == CHARSXP:
            n = self.i32()
            return None if n == -1 else self.raw(n).decode("utf-8", "replace")
        if t in (LGLSXP, INTSXP):
            n = self.i32()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif t == REALSXP:
            n = self.i32()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif t == STRSXP:
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == VECSXP:
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t} at byte {self.p}")
        attr = {}
        if has_attr:
            for tag, car in self.item():
                attr[tag] = car
        return RObj(v, attr)


def read_rda(path):
    """Return {object name: RObj} for an R `save()` file."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:3] == b"BZh":
        buf = bz2.decompress(buf)
    elif buf[:2] == b"\x1f\x8b":
        buf = gzip.decompress(buf)
    elif buf[:6] == b"\xfd7zXZ\x00":
        buf = lzma.decompress(buf)
    if buf[:5] != b"RDX2\n" or buf[5:7] != b"X\n":
        raise ValueError("not an RDX2/XDR file")
    r = _Reader(buf)
    r.p = 7
    r.i32(), r.i32(), r.i32()
    return {tag: car for tag, car in r.item()}


def names(o):
    return list(o.attr["names"].value) if "names" in o.attr else None


def as_factor(o):
    """INTSXP + levels -> (codes 1-based int32 array, list of level strings)."""
    return o.value, list(o.attr["levels"].value)


def data_frame_columns(o):
    """VECSXP with class data.frame -> ordered {column name: RObj}."""
    return dict(zip(names(o), o.value))
