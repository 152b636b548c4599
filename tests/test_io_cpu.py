"""tlab restart / grid file formats (tlab_b200/io.py) against the byte layout of src/base/tlab_grid.f90:26-91 and
src/base/io_fields.f90:150-264,346-456,534-596 (stream access)."""
import struct

import numpy as np
import pytest

from tlab_b200 import io as tio


def test_grid_file_is_fortran_sequential_unformatted(tmp_path):
    x, y, z = np.arange(4) * 0.5, np.array([0.0, 0.1, 0.3]), np.array([0.0])
    name = str(tmp_path / "grid")
    tio.grid_write(name, x, y, z, scales=(2.0, 0.3, 1.0))
    raw = open(name, "rb").read()
    # record 1: three int32 framed by their byte count (12)
    assert struct.unpack("<i3ii", raw[:20]) == (12, 4, 3, 1, 12)
    # record 2: three fp64 framed by 24
    assert struct.unpack("<i3di", raw[20:52]) == (24, 2.0, 0.3, 1.0, 24)
    # record 3: x nodes framed by 32
    assert struct.unpack("<i", raw[52:56])[0] == 32
    assert np.array_equal(np.frombuffer(raw[56:88], dtype="<f8"), x)
    assert len(raw) == 20 + 32 + (8 + 32) + (8 + 24) + (8 + 8)
    x2, y2, z2, sc = tio.grid_read(name, sizes=(4, 3, 1))
    assert np.array_equal(x2, x) and np.array_equal(y2, y) and np.array_equal(z2, z) and sc == (2.0, 0.3, 1.0)
    with pytest.raises(tio.TlabIOError) as e:
        tio.grid_read(name, sizes=(4, 3, 2))
    assert e.value.code == tio.DNS_ERROR_DIMGRID


def test_field_file_header_and_payload(tmp_path):
    nx, ny, nz = 5, 3, 4
    rng = np.random.default_rng(3)
    q = [rng.standard_normal((nz, ny, nx)) for _ in range(3)]
    fname = str(tmp_path / "flow.10")
    tio.write_fields(fname, 10, q, tio.flow_params(0.25, 2e-4))
    raw = open(fname + ".2", "rb").read()
    assert struct.unpack("<5i", raw[:20]) == (20 + 4 * 8, nx, ny, nz, 10)          # offset, nx, ny, nz, nt
    assert struct.unpack("<4d", raw[20:52]) == (0.25, 2e-4, 1.0, 1.0)              # rtime, visc, froude, rossby
    assert len(raw) == 52 + nx * ny * nz * 8
    # x fastest: element (i, j, k) of a(nx, ny, nz) at i + nx*(j + ny*k)
    i, j, k = 3, 1, 2
    assert struct.unpack("<d", raw[52 + 8 * (i + nx * (j + ny * k)):][:8])[0] == q[1][k, j, i]
    back, nt, params = tio.read_fields(fname, nx, ny, nz, 3)
    assert nt == 10 and list(params) == [0.25, 2e-4, 1.0, 1.0]
    for a, b in zip(back, q):
        assert np.array_equal(a, b)
    one, _, _ = tio.read_fields(fname, nx, ny, nz, 3, iread=3)
    assert len(one) == 1 and np.array_equal(one[0], q[2])
    with pytest.raises(tio.TlabIOError) as e:
        tio.read_fields(fname, nx, ny + 1, nz, 3)
    assert e.value.code == tio.DNS_ERROR_DIMGRID


def test_scalar_headers_per_field_and_slab_access(tmp_path):
    """One header per scalar (rtime, visc, schmidt(is)); z-slabs written and read in place like the MPI-IO sub-array view."""
    nx, ny, nz, P = 4, 3, 8, 2
    rng = np.random.default_rng(4)
    s = [rng.standard_normal((nz, ny, nx)) for _ in range(2)]
    fname = str(tmp_path / "scal.3")
    headers = [tio.scal_params(1.5, 1e-3, 1.0), tio.scal_params(1.5, 1e-3, 0.7)]
    kmax = nz // P
    for r in range(P):
        tio.write_fields(fname, 3, [a[r * kmax:(r + 1) * kmax] for a in s], headers, koff=r * kmax, nz_total=nz)
    _, _, p2 = tio.read_fields(fname, nx, ny, nz, 2, iread=2)
    assert list(p2) == [1.5, 1e-3, 0.7]
    full, _, _ = tio.read_fields(fname, nx, ny, nz, 2)
    for a, b in zip(full, s):
        assert np.array_equal(a, b)
    for r in range(P):
        slab, _, _ = tio.read_fields(fname, nx, ny, nz, 2, koff=r * kmax, kmax=kmax)
        assert np.array_equal(slab[1], s[1][r * kmax:(r + 1) * kmax])


def test_broken_header_is_rejected(tmp_path):
    name = str(tmp_path / "flow.1")
    with open(name + ".1", "wb") as f:
        f.write(struct.pack("<5i", 23, 2, 2, 2, 1))       # offset not 20 + 8*k
        f.write(b"\0" * 100)
    with pytest.raises(tio.TlabIOError) as e:
        tio.read_fields(name, 2, 2, 2, 1)
    assert e.value.code == tio.DNS_ERROR_RECLEN


def test_slab_writes_need_no_ordering_between_ranks(tmp_path):
    """Slab-wise IO_Write_Fields: the header rank may come last (or find a stale, longer file) -- the file is the same."""
    from tlab_b200 import io
    rng = np.random.default_rng(5)
    nx, ny, nz = 8, 5, 12
    a = rng.standard_normal((nz, ny, nx))
    ref, out = str(tmp_path / "ref"), str(tmp_path / "out")
    io.write_fields(ref, 7, [a], params=[0.5, 1e-3])
    with open(io.field_name(out, 1), "wb") as f:
        f.write(b"\xff" * (3 * a.nbytes))                     # a stale longer file
    for koff in (8, 4, 0):                                      # last slab first, header rank last
        io.write_fields(out, 7, [a[koff:koff + 4]], params=[0.5, 1e-3], koff=koff, nz_total=nz)
    assert open(io.field_name(out, 1), "rb").read() == open(io.field_name(ref, 1), "rb").read()
