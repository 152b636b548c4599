"""Domain decomposition plumbing: z slabs, one process per GPU.

Mirrors the reference's rank layout (src/base/tlab_mpi_procs.f90:40-58, ims_npro_k = world size, ims_npro_i = 1)
and the K-transpose index maps (src/base/tlab_mpi_transpose.f90:301-325).  torch.distributed is used only to
exchange the NCCL unique id (the Fortran host would MPI_Bcast it); the transposes of the product run inside
libtlab_gpu.so on its own NCCL communicator.

(The torch restatement of the pack/unpack layout used by the gloo CPU test lives in tests/trp_layout_ref.py.)
"""
import ctypes

import torch

from . import lib as _lib


def init_from_torch_distributed():
    """Create the library's NCCL communicator over the default torch.distributed process group."""
    import torch.distributed as dist
    L = _lib.load()
    rank, world = dist.get_rank(), dist.get_world_size()
    idbuf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0 and world > 1:
        raw = (ctypes.c_ubyte * 128)()
        _lib.check(L.tlab_mpi_get_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = idbuf.to(dev)
        dist.broadcast(t, src=0)
        idbuf = t.cpu()
    raw = (ctypes.c_ubyte * 128)(*[int(v) for v in idbuf.tolist()])
    _lib.check(L.tlab_mpi_init(rank, world, raw))
    return rank, world


def finalize():
    _lib.load().tlab_mpi_finalize()


def slab(nz, rank, world):
    """(kmax, offset) of this rank's z slab; the reference requires nz % world == 0 (tlab_mpi_procs.f90:44-58)."""
    if nz % world:
        raise ValueError("Kmax(*) must divide the grid size in z")
    kmax = nz // world
    return kmax, rank * kmax
