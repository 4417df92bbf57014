"""Minimal reader for R's `.rda` (xz -> RDX2 XDR serialization), enough for data.frames and
sf geometry lists.  Used ONLY by make_golden.py in the build container
to turn /root/reference/data/*.rda into small committed fixtures; nothing reads R files at
test time."""
from __future__ import annotations

import lzma
import struct

import numpy as np


class _R:
    def __init__(self, buf):
        self.b, self.p, self.refs = buf, 0, []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:                      # NILVALUE
            return None
        if t == 255:                      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:                        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:                        # LISTSXP (pairlist)
            out = {}
            while True:
                attr = self.item() if has_attr else None   # noqa: F841
                tag = self.item() if has_tag else None
                out[tag if tag is not None else len(out)] = self.item()
                flags = self.i32()
                t = flags & 0xFF
                if t == 254:
                    return out
                assert t == 2, t
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        if t == 9:                        # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.p:self.p + n].decode("utf-8", "replace")
            self.p += n
            return s
        if t in (10, 13):                 # LGLSXP, INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, ">i4", n, self.p).astype(np.int64)
            self.p += 4 * n
        elif t == 14:                     # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, ">f8", n, self.p).astype(np.float64)
            self.p += 8 * n
        elif t == 16:                     # STRSXP
            v = [self.item() for _ in range(self.i32())]
        elif t == 19:                     # VECSXP
            v = [self.item() for _ in range(self.i32())]
        else:
            raise NotImplementedError(f"SEXP type {t}")
        attrs = self.item() if has_attr else None
        if attrs:
            if isinstance(v, np.ndarray) and "dim" in attrs:
                v = v.reshape(tuple(int(d) for d in attrs["dim"]), order="F")
            if isinstance(v, list) and "names" in attrs and t == 19:
                return {"__names__": attrs["names"], "__values__": v, "__attrs__": attrs}
        return v


def read_rda(path):
    buf = lzma.open(path).read()
    assert buf[:7] == b"RDX2\nX\n", buf[:7]
    r = _R(buf)
    r.p = 7
    r.i32(); r.i32(); r.i32()            # format version, writer version, min reader version
    return r.item()                      # pairlist {object name: value}


def data_frame(obj):
    return dict(zip(obj["__names__"], obj["__values__"]))
