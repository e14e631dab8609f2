"""Reader / writer of Idefix restart dumps (``dump.NNNN.dmp``).

Format (src/output/dump.cpp:812-985, reader counterpart pytools/dump_io.py:30-124):
128-byte header ``Idefix <ver> Dump Data little endian``; then fields
``{char name[16]; int32 type; int32 ndim; int32 dims[ndim]; raw data}`` with
type 0 = double, 1 = float, 2 = int32, 3 = bool: ``x1,xl1,xr1, ... xr3`` (interior cells),
every registered variable in alphabetical order (``Vc-*`` interior cells, ``Vs-*`` with one
more face in their own direction; cell-centred ``BX*`` are not dumped in MHD,
fluid.hpp:607-622), the scalars, and a terminating ``eof`` field.
Arrays are stored i-fastest; they are returned indexed [i, j, k] like dump_io does.
"""
import re
import struct

import numpy as np

NAME_SIZE = 16
HEADER_SIZE = 128
_TYPES = {0: ("<f8", 8), 1: ("<f4", 4), 2: ("<i4", 4), 3: ("?", 1)}
_HEADER_RE = re.compile(r"Idefix (?P<version>[\w\.-]+) Dump Data((?P<byteorder>(little|big)) endian)?")


class DumpDataset:
    """x1, xl1, xr1, ... plus ``data[name]`` for every other field."""

    def __init__(self, filename=None):
        self.filename = filename
        self.header = ""
        self.metadata = {}
        self.data = {}
        if filename is not None:
            with open(filename, "rb") as fh:
                self._read(fh)

    def _read_field(self, fh):
        raw = fh.read(NAME_SIZE)
        name = raw[:raw.index(b"\x00")].decode()
        typ, ndim = struct.unpack("<ii", fh.read(8))
        dims = struct.unpack(f"<{ndim}i", fh.read(4 * ndim))
        dtype, size = _TYPES[typ]
        n = int(np.prod(dims))
        arr = np.frombuffer(fh.read(size * n), dtype=dtype).reshape(dims[::-1]).T
        return name, arr

    def _read(self, fh):
        raw = fh.read(HEADER_SIZE)
        self.header = raw[:raw.index(b"\x00")].decode()
        m = _HEADER_RE.match(self.header)
        if m is None:
            raise ValueError(f"{self.filename}: not an Idefix dump file")
        self.metadata = {"version": m.group("version"), "byteorder": m.group("byteorder")}
        for attr in ("x1", "x1l", "x1r", "x2", "x2l", "x2r", "x3", "x3l", "x3r"):
            setattr(self, attr, self._read_field(fh)[1])
        while True:
            name, arr = self._read_field(fh)
            if name == "eof":
                break
            self.data[name] = arr


def readDump(filename):
    return DumpDataset(filename)


def _write_field(fh, name, arr, typ):
    dtype = _TYPES[typ][0]
    a = np.asarray(arr, dtype=dtype)
    dims = a.shape if a.ndim else (1,)
    fh.write(name.encode()[:NAME_SIZE - 1].ljust(NAME_SIZE, b"\x00"))
    fh.write(struct.pack("<ii", typ, len(dims)))
    fh.write(struct.pack(f"<{len(dims)}i", *dims))
    # arrays are handed over indexed [i, j, k]; the file is i-fastest
    fh.write(np.ascontiguousarray(a.T if a.ndim > 1 else a).tobytes())


def writeDump(filename, data, vc_names, Vc, Vs=None, t=0.0, dt=0.0, number=0, version="2.2.02-b200"):
    """Write the interior of Vc (nvar,k,j,i; only the entries named in vc_names, skipping None)
    and Vs (ndim,k,j,i) of DataBlock ``data`` in the reference's dump format (``number`` = N of dump.N.dmp)."""
    grid = data.mygrid
    nd = data.defs.dimensions
    with open(filename, "wb") as fh:
        head = f"Idefix {version} Dump Data little endian".encode()
        fh.write(head.ljust(HEADER_SIZE, b"\x00"))
        for d in range(3):
            g = grid.nghost[d]
            sl = slice(g, g + grid.np_int[d])
            _write_field(fh, f"x{d + 1}", grid.x[d][sl], 0)
            _write_field(fh, f"xl{d + 1}", grid.xl[d][sl], 0)
            _write_field(fh, f"xr{d + 1}", grid.xr[d][sl], 0)
        kb, ke, jb, je, ib, ie = data.beg[2], data.end[2], data.beg[1], data.end[1], data.beg[0], data.end[0]
        fields = {}
        for n, name in enumerate(vc_names):
            if name is not None:
                fields[f"Vc-{name}"] = Vc[n, kb:ke, jb:je, ib:ie].transpose(2, 1, 0)
        if Vs is not None:
            for d in range(nd):
                o = [1 if e == d else 0 for e in range(3)]
                fields[f"Vs-BX{d + 1}s"] = Vs[d, kb:ke + o[2], jb:je + o[1], ib:ie + o[0]].transpose(2, 1, 0)
        # dump.N stores N+1: the reference increments the counter before it writes the fields (dump.cpp:839)
        scalars = {"time": (t, 0), "dt": (dt, 0), "dumpFileNumber": (number + 1, 2),
                   "geometry": (data.defs.geometry, 2),
                   "periodicity": ([int(grid.lbound[d] == 1) for d in range(3)], 2)}
        for name in sorted(list(fields) + list(scalars)):
            if name in fields:
                _write_field(fh, name, fields[name], 0)
            else:
                val, typ = scalars[name]
                _write_field(fh, name, np.atleast_1d(val), typ)
        _write_field(fh, "eof", np.zeros(1), 0)


def vc_names(defs, ntracer=0):
    """names of the dumped Vc entries in variable order; None for those not dumped"""
    names = ["RHO"] + [f"VX{c + 1}" for c in range(defs.components)]
    if not defs.isothermal:
        names.append("PRS")
    if defs.mhd:
        # components evolved by CT are dumped through Vs only; BX(n > DIMENSIONS) stays in Vc
        names += [None if c < defs.dimensions else f"BX{c + 1}" for c in range(defs.components)]
    names += [f"TR{n}" for n in range(ntracer)]   # fluid.hpp:634
    return names


def compare(a, b, fields=None):
    """reference metric (pytools/idfx_test.py:414-428): mean over common fields of RMS difference"""
    errs = {}
    for name in a.data:
        if name in b.data and name.startswith(("Vc-", "Vs-")) and (fields is None or name in fields):
            x, y = np.asarray(a.data[name], float), np.asarray(b.data[name], float)
            if x.shape == y.shape:
                errs[name] = float(np.sqrt(np.mean((x - y) ** 2)))
    return (sum(errs.values()) / max(len(errs), 1)), errs
