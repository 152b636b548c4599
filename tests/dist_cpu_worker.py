"""gloo worker (CPU, world_size 2): the K-transpose layout (tests/trp_layout_ref.py) against the global-array definition, and the
slab-wise restart files of a split domain."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import trp_layout_ref as R  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from tlab_b200 import mpi
    nxy, nz = 24, 8
    kmax, koff = mpi.slab(nz, rank, world)
    rng = np.random.default_rng(5)
    full = rng.standard_normal((nz, nxy))                 # Fortran a(nxy, nz) = C (nz, nxy)
    a = torch.from_numpy(full[koff:koff + kmax].copy())
    b = R.trp_k_forward_ref(a)                          # pencil (nz, nxy/P): lines rank*nl .. (rank+1)*nl
    nl = nxy // world
    expect = full[:, rank * nl:(rank + 1) * nl]
    assert np.array_equal(b.numpy(), expect), "forward map (tlab_mpi_transpose.f90:301-325)"
    back = R.trp_k_backward_ref(b, kmax)
    assert torch.equal(back, a), "backward is the inverse"
    # complex data: (re, im) pairs travel together
    cf = rng.standard_normal((nz, nxy)) + 1j * rng.standard_normal((nz, nxy))
    ca = torch.from_numpy(np.ascontiguousarray(cf[koff:koff + kmax]).view(np.float64).copy())   # (kmax, 2*nxy)
    # treat complex elements as blocks of 2 doubles: partition by complex lines
    cb = R.trp_k_forward_ref(ca.reshape(kmax, nxy, 2).reshape(kmax, nxy * 2))
    got = cb.numpy().reshape(nz, nl, 2)
    exp = np.ascontiguousarray(cf[:, rank * nl:(rank + 1) * nl]).view(np.float64).reshape(nz, nl, 2)
    assert np.array_equal(got, exp)
    # pack/unpack are inverse permutations
    assert torch.equal(R.unpack_k(R.pack_k(a, world)), a)
    # restart files of a split domain (tlab_b200/io.py): every rank writes its z-slab in place, like the MPI-IO sub-array
    # view of IO_Write_Fields (io_fields.f90:346-456); rank 0 owns the header; everybody reads its slab back
    from tlab_b200 import io as tio
    nx, ny = 6, 4
    fields = [rng.standard_normal((nz, ny, nx)) for _ in range(2)]
    fname = os.path.join(os.environ.get("TLAB_TMP", "/tmp"), "flow_gloo.7")
    mine = [f[koff:koff + kmax] for f in fields]
    for turn in range(world):                             # the rank holding plane 0 creates the files first
        if turn == rank:
            tio.write_fields(fname, 7, mine, tio.flow_params(0.5, 1e-3), koff=koff, nz_total=nz)
        dist.barrier()
    back_slab, nt, params = tio.read_fields(fname, nx, ny, nz, 2, koff=koff, kmax=kmax)
    assert nt == 7 and list(params) == [0.5, 1e-3, 1.0, 1.0]
    for got_f, exp_f in zip(back_slab, mine):
        assert np.array_equal(got_f, exp_f)
    whole, _, _ = tio.read_fields(fname, nx, ny, nz, 2)
    for got_f, exp_f in zip(whole, fields):
        assert np.array_equal(got_f, exp_f)
    dist.barrier()
    if rank == 0:
        for i in (1, 2):
            os.remove(tio.field_name(fname, i))
        print("DIST_CPU_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
