"""tlab's restart and grid file formats, so that the GPU path can start from and hand back
real tlab checkpoints.

* grid file (`TLab_Grid_Read` / `TLab_Grid_Write`, src/base/tlab_grid.f90:26-91): Fortran *sequential unformatted*
  records, each framed by two int32 byte counts: `(nx, ny, nz)` int32, `(scalex, scaley, scalez)` fp64, `x(:)`, `y(:)`,
  `z(:)` fp64.
* field files (`IO_Read_Fields` / `IO_Write_Fields`, src/base/io_fields.f90:150-264,346-456; header
  `IO_READ_HEADER` / `IO_WRITE_HEADER`, :534-596), stream access (`USE_ACCESS_STREAM`, src/include/types.h:14): one
  file per field, `<name>.<ifield>`; int32 `(offset, nx, ny, nz, nt)` with `offset = 20 + 8*len(params)`, then `params`
  fp64, then the raw fp64 array `a(nx, ny, nz)`, x fastest, starting at byte `offset`.  `nx, ny, nz` are the GLOBAL
  extents; an MPI rank reads its block through a sub-array view (`IO_Create_Subarray_XOZ`), here a z-slab
  `[koff, koff + kmax)`.  Header parameters of the flow files: rtime, visc, froude, rossby [, gama0, prandtl, mach];
  of scalar file `is`: rtime, visc, schmidt(is)  (src/physics/tlab_consistency_check.f90:148-163).

Arrays are returned / taken in C order `(nz, ny, nx)`, the layout of the rest of this package.  Pure host code (numpy).
"""
import os
import struct

import numpy as np

SIZEOFINT = 4
SIZEOFREAL = 8
DNS_ERROR_DIMGRID = 48     # src/include/dns_error.h:43
DNS_ERROR_RECLEN = 43      # src/include/dns_error.h:38


class TlabIOError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (error %d)" % (msg, code))
        self.code = code


# ---------------------------------------------------------------------------------------------- grid
def _record(payload):
    n = struct.pack("<i", len(payload))
    return n + payload + n


def grid_write(name, x, y, z, scales=None):
    """TLab_Grid_Write.  `scales` = (scalex, scaley, scalez); default: the extent of each node array, with one more
    spacing for a uniform periodic direction being the caller's business (the reference stores what grid.ini gave)."""
    x, y, z = (np.ascontiguousarray(a, dtype="<f8") for a in (x, y, z))
    if scales is None:
        scales = [float(a[-1] - a[0]) if a.size > 1 else 1.0 for a in (x, y, z)]
    with open(name, "wb") as f:
        f.write(_record(struct.pack("<3i", x.size, y.size, z.size)))
        f.write(_record(struct.pack("<3d", *[float(s) for s in scales])))
        for a in (x, y, z):
            f.write(_record(a.tobytes()))


def _read_record(f, what):
    head = f.read(4)
    if len(head) != 4:
        raise TlabIOError(DNS_ERROR_RECLEN, "grid file: missing record (%s)" % what)
    (n,) = struct.unpack("<i", head)
    payload = f.read(n)
    tail = f.read(4)
    if len(payload) != n or len(tail) != 4 or struct.unpack("<i", tail)[0] != n:
        raise TlabIOError(DNS_ERROR_RECLEN, "grid file: broken record framing (%s)" % what)
    return payload


def grid_read(name, sizes=None):
    """TLab_Grid_Read -> (x, y, z, scales).  `sizes`, when given, must match the file (DNS_ERROR_DIMGRID otherwise)."""
    with open(name, "rb") as f:
        n = struct.unpack("<3i", _read_record(f, "sizes"))
        if sizes is not None and tuple(int(s) for s in sizes) != n:
            raise TlabIOError(DNS_ERROR_DIMGRID, "grid file: dimensions (%d,%d,%d) unmatched" % n)
        scales = struct.unpack("<3d", _read_record(f, "scales"))
        out = []
        for i, nm in enumerate("xyz"):
            a = np.frombuffer(_read_record(f, nm), dtype="<f8")
            if a.size != n[i]:
                raise TlabIOError(DNS_ERROR_RECLEN, "grid file: %s has %d nodes, header says %d" % (nm, a.size, n[i]))
            out.append(a.copy())
    return out[0], out[1], out[2], scales


# ---------------------------------------------------------------------------------------------- fields
def field_name(fname, ifield):
    """`flow.<it>` + field index -> `flow.<it>.<ifield>` (io_fields.f90:208-209: write(name,'(I2)') ifield, left-adjusted)."""
    return "%s.%d" % (fname, ifield)


def write_header(f, nx, ny, nz, nt, params=()):
    params = np.asarray(params, dtype="<f8").ravel()
    offset = 5 * SIZEOFINT + params.size * SIZEOFREAL
    f.write(struct.pack("<5i", offset, nx, ny, nz, nt))
    f.write(params.tobytes())
    return offset


def read_header(f, nx=None, ny=None, nz=None):
    """IO_READ_HEADER -> (offset, (nx, ny, nz), nt, params).  Extents, when given, are checked (DNS_ERROR_DIMGRID)."""
    raw = f.read(5 * SIZEOFINT)
    if len(raw) != 5 * SIZEOFINT:
        raise TlabIOError(DNS_ERROR_RECLEN, "field file: header too short")
    offset, fx, fy, fz, nt = struct.unpack("<5i", raw)
    for want, got in ((nx, fx), (ny, fy), (nz, fz)):
        if want is not None and want != got:
            raise TlabIOError(DNS_ERROR_DIMGRID, "IO_READ_HEADER. Grid size mismatch.")
    isize = offset - 5 * SIZEOFINT
    if isize < 0 or isize % SIZEOFREAL:
        raise TlabIOError(DNS_ERROR_RECLEN, "IO_READ_HEADER. Header format incorrect.")
    params = np.frombuffer(f.read(isize), dtype="<f8").copy()
    if params.size * SIZEOFREAL != isize:
        raise TlabIOError(DNS_ERROR_RECLEN, "IO_READ_HEADER. Header format incorrect.")
    return offset, (fx, fy, fz), nt, params


def write_fields(fname, nt, fields, params=(), koff=0, nz_total=None):
    """IO_Write_Fields: fields[i] (nz_local, ny, nx) -> `<fname>.<i+1>`.  `params` is one list for all fields or one list
    per field (`locHeader(min(size, ifield))`).  With `nz_total` the arrays are z-slabs starting at plane `koff` of a
    global field: the header is written by the rank with koff = 0 and every rank writes its planes in place."""
    per_field = len(params) > 0 and np.ndim(params[0]) > 0
    for i, a in enumerate(fields):
        a = np.ascontiguousarray(a, dtype="<f8")
        nzl, ny, nx = a.shape
        nzt = nzl if nz_total is None else int(nz_total)
        p = params[min(i, len(params) - 1)] if per_field else params
        name = field_name(fname, i + 1)
        # Every rank opens the file create-if-missing and WITHOUT truncation, and every byte of it is written by exactly one
        # rank (the header by the rank with koff = 0, each slab by its owner), so the ranks need no ordering among themselves
        # (the reference puts an MPI_BARRIER between its header write and the MPI-IO data write; src/io/io_fields.f90).
        # A stale longer file is cut to size by the header rank.
        offset = 5 * SIZEOFINT + len(np.ravel(p)) * SIZEOFREAL
        fd = os.open(name, os.O_RDWR | os.O_CREAT, 0o644)
        with os.fdopen(fd, "r+b") as f:
            if koff == 0:
                assert write_header(f, nx, ny, nzt, nt, p) == offset
                f.truncate(offset + nx * ny * nzt * SIZEOFREAL)
            f.seek(offset + koff * ny * nx * SIZEOFREAL)
            f.write(a.tobytes())


def read_fields(fname, nx, ny, nz, nfield, iread=0, koff=0, kmax=None):
    """IO_Read_Fields -> (list of (kmax, ny, nx) arrays, nt, params of the last header read).  `nx, ny, nz` are the global
    extents; `iread` = 0 reads fields 1..nfield, otherwise that one field; `koff, kmax` select a z-slab."""
    kmax = nz - koff if kmax is None else kmax
    out, nt, params = [], None, None
    for ifield in range(1, nfield + 1):
        if iread not in (0, ifield):
            continue
        with open(field_name(fname, ifield), "rb") as f:
            offset, _, nt, params = read_header(f, nx, ny, nz)
            need = offset + nx * ny * nz * SIZEOFREAL
            if os.fstat(f.fileno()).st_size < need:
                raise TlabIOError(DNS_ERROR_RECLEN, "field file shorter than its header says")
            f.seek(offset + koff * ny * nx * SIZEOFREAL)
            a = np.fromfile(f, dtype="<f8", count=kmax * ny * nx)
        out.append(a.reshape(kmax, ny, nx))
    return out, nt, params


def flow_params(rtime, visc, froude=1.0, rossby=1.0):
    return [rtime, visc, froude, rossby]


def scal_params(rtime, visc, schmidt):
    return [rtime, visc, schmidt]
